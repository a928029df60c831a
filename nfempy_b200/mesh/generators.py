"""Synthetic simplex meshes laid out the way a gmsh file would be (boundary blocks first, domain last).

gmsh is absent, so the benchmark workloads of SURVEY.md section 8(d) are built here: a unit cube cut into
``n^3`` cubes x 6 Kuhn tetrahedra with interior vertices jittered, and a unit square cut into ``n^2 x 2``
right triangles.  Physical tags follow the reference fixtures
(R/examples/cosserat_elasticity/gmsh_files/ex_3/example_3_3d_l_0.msh: domain 1, sides 2..7).
"""
import itertools

import numpy as np

from .mesh import Mesh, TET_FACES, TRI_EDGES


def box_tet_points_cells(n, jitter=0.2, seed=1234, shuffle_local_nodes=False, lengths=(1.0, 1.0, 1.0), nz=None,
                         nx=None):
    """Points and tets ``[6 nx n nz, 4]`` of a Kuhn-subdivided box of ``nx x n x nz`` cubes of edge ``1/n``
    (``nx``, ``nz`` default to n).  Cells are cube-major with x slowest, so contiguous cell ranges are x-slabs."""
    nz = n if nz is None else nz
    nx = n if nx is None else nx
    rng = np.random.default_rng(seed)
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(n + 1), np.arange(nz + 1), indexing="ij")
    grid = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    dims = np.array([nx, n, nz])
    h = np.asarray(lengths, dtype=np.float64) / np.array([n, n, n])
    points = grid * h
    if jitter > 0.0:
        interior = np.all((grid > 0) & (grid < dims), axis=1)
        shift = rng.uniform(-jitter, jitter, size=points.shape) * h
        points[interior] += shift[interior]

    def nid(i, j, k):
        return (i * (n + 1) + j) * (nz + 1) + k

    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(n), np.arange(nz), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
    tets = []
    for perm in itertools.permutations(range(3)):
        corner = [np.stack([ci, cj, ck], axis=1)]
        for axis in perm:
            step = np.zeros(3, dtype=np.int64)
            step[axis] = 1
            corner.append(corner[-1] + step)
        tets.append(np.stack([nid(c[:, 0], c[:, 1], c[:, 2]) for c in corner], axis=1))
    # cube-major order: the 6 tets of a cube are consecutive
    tets = np.stack(tets, axis=1).reshape(-1, 4)
    if shuffle_local_nodes:
        keys = rng.random(tets.shape)
        tets = np.take_along_axis(tets, np.argsort(keys, axis=1), axis=1)
    # gmsh emits positively oriented tetrahedra; the reference relies on it (its Piola map uses |det J|,
    # R/src/geometry/mapping.py:28-32), so do the same: swap the last two nodes of left-handed tets
    X = points[tets]
    det = np.einsum("ni,ni->n", X[:, 1] - X[:, 0], np.cross(X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]))
    neg = det < 0
    tets[neg] = tets[neg][:, [0, 1, 3, 2]]
    return points, tets.astype(np.int64), grid, dims


def box_tet_mesh(n, jitter=0.2, seed=1234, shuffle_local_nodes=False, nz=None, lengths=(1.0, 1.0, 1.0), nx=None):
    """``Mesh`` of the Kuhn box: six boundary triangle blocks (tags 2..7), then the tets (tag 1)."""
    points, tets, grid, dims = box_tet_points_cells(n, jitter, seed, shuffle_local_nodes, lengths, nz, nx)
    faces = tets[:, TET_FACES].reshape(-1, 3)
    fg = grid[faces]  # [nf, 3 nodes, 3 coords]
    blocks = []
    # south y=0 (2), east x=n (3), north y=n (4), west x=0 (5), top z=nz (6), bottom z=0 (7)
    sides = [(2, 1, 0), (3, 0, dims[0]), (4, 1, dims[1]), (5, 0, 0), (6, 2, dims[2]), (7, 2, 0)]
    for tag, axis, value in sides:
        on = np.all(fg[:, :, axis] == value, axis=1)
        tri = faces[on]
        blocks.append(("triangle", tri, np.full(len(tri), tag, dtype=np.int64)))
    blocks.append(("tetra", tets, np.ones(len(tets), dtype=np.int64)))
    mesh = Mesh.from_arrays(3, points, blocks)
    mesh.build_conformal_mesh()
    return mesh


def square_tri_mesh(n, jitter=0.0, seed=1234):
    """Unit square, ``n x n x 2`` right triangles; boundary line blocks tags 2..5, triangles tag 1."""
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    grid = np.stack([gx.ravel(), gy.ravel()], axis=1)
    h = 1.0 / n
    points = np.zeros((len(grid), 3))
    points[:, :2] = grid * h
    if jitter > 0.0:
        interior = np.all((grid > 0) & (grid < n), axis=1)
        shift = rng.uniform(-jitter, jitter, size=(len(grid), 2)) * h
        points[interior, :2] += shift[interior]

    def nid(i, j):
        return i * (n + 1) + j

    ci, cj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    ci, cj = ci.ravel(), cj.ravel()
    lower = np.stack([nid(ci, cj), nid(ci + 1, cj), nid(ci + 1, cj + 1)], axis=1)  # counter-clockwise, like gmsh
    upper = np.stack([nid(ci, cj), nid(ci + 1, cj + 1), nid(ci, cj + 1)], axis=1)
    tris = np.stack([lower, upper], axis=1).reshape(-1, 3).astype(np.int64)
    edges = tris[:, TRI_EDGES].reshape(-1, 2)
    eg = grid[edges]
    blocks = []
    for tag, axis, value in [(2, 1, 0), (3, 0, n), (4, 1, n), (5, 0, 0)]:
        on = np.all(eg[:, :, axis] == value, axis=1)
        blocks.append(("line", edges[on], np.full(int(on.sum()), tag, dtype=np.int64)))
    blocks.append(("triangle", tris, np.ones(len(tris), dtype=np.int64)))
    mesh = Mesh.from_arrays(2, points, blocks)
    mesh.build_conformal_mesh()
    return mesh


def fractured_box_tet_mesh(n, jitter=0.0, seed=1234, shuffle_local_nodes=False):
    """Unit cube of ``n^3`` cubes x 6 Kuhn tetrahedra (``n`` even) cut by the horizontal fracture plane ``z = 1/2`` -- the
    mixed-dimensional mesh of BASELINE config 5 ("fractured 3D mesh"), laid out the way the reference leaves a mesh after
    ``cut_conformity_along_c1_lines`` (R/src/mesh/mesh_operations.py:274-398, which only handles 2-D): the fracture keeps
    its own cells on the ORIGINAL nodes (triangles, tag 10; their boundary edges on the box sides, tag 30), and every
    node of the plane is duplicated twice so that the tetrahedra above (``+``) and below (``-``) the plane no longer share
    it; the faces of those tetrahedra on the plane are boundary facets of the 3-D domain (tag 50, "clones"), listed with
    the same vertex order as the fracture triangle they face.  Blocks: six outer sides (tags 2..7), + clones, - clones,
    fracture boundary edges, fracture triangles, tetrahedra (tag 1).

    Returns ``(mesh, interface)`` with ``interface = {"c1": tri[m, 3], "positive": tri[m, 3], "negative": tri[m, 3]}``
    (node tags of the fracture triangle and of its two clones, row by row) -- the triples the coupling form is
    scattered over (R/examples/md_elliptic_scalar/md_elliptic_scalar.py:324-342).  The reference performs the cut in
    place (clones get fresh cell ids appended to the mesh); here the cut mesh is numbered from scratch by
    ``Mesh.build_conformal_mesh``, a different but equivalent numbering of the same cells."""
    if n % 2:
        raise ValueError("fractured_box_tet_mesh: n must be even (the fracture is the grid plane z = 1/2)")
    points, tets, grid, dims = box_tet_points_cells(n, jitter, seed, shuffle_local_nodes)
    k0 = n // 2
    on_plane = grid[:, 2] == k0
    points[on_plane, 2] = 0.5  # the plane stays flat under jitter
    n0 = len(points)
    plane_nodes = np.flatnonzero(on_plane)
    slot = np.full(n0, -1, dtype=np.int64)
    slot[plane_nodes] = np.arange(len(plane_nodes))
    pos_of = np.arange(n0)
    neg_of = np.arange(n0)
    pos_of[plane_nodes] = n0 + slot[plane_nodes]
    neg_of[plane_nodes] = n0 + len(plane_nodes) + slot[plane_nodes]
    points = np.concatenate([points, points[plane_nodes], points[plane_nodes]], axis=0)
    grid = np.concatenate([grid, grid[plane_nodes], grid[plane_nodes]], axis=0)
    # a tet lies above the plane iff its lowest vertex does
    above = grid[tets].min(axis=1)[:, 2] >= k0
    # fracture triangles: faces of the tets above that lie in the plane, in that tet's local face order
    faces = tets[above][:, TET_FACES].reshape(-1, 3)
    c1 = faces[np.all(on_plane[faces], axis=1)]
    # consistently oriented fracture triangles (normal +z), as a mesher leaves the cells of a plane surface: the 2-D
    # orientation rule of the H(div) edge functions (R/src/basis/permute_and_transform.py:8-19) only gives a continuous
    # normal flux across an edge when both triangles turn the same way
    P = points[c1]
    flip = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])[:, 2] < 0.0
    c1[flip] = c1[flip][:, [0, 2, 1]]
    new_tets = tets.copy()
    new_tets[above] = pos_of[tets[above]]
    new_tets[~above] = neg_of[tets[~above]]
    positive, negative = pos_of[c1], neg_of[c1]
    # boundary edges of the fracture: edges of its triangles on the four vertical sides of the box
    e = c1[:, TRI_EDGES].reshape(-1, 2)
    eg = grid[e]
    side = np.zeros(len(e), dtype=bool)
    for axis, value in ((0, 0), (0, dims[0]), (1, 0), (1, dims[1])):
        side |= np.all(eg[:, :, axis] == value, axis=1)
    c1_boundary = e[side]
    all_faces = new_tets[:, TET_FACES].reshape(-1, 3)
    fg = grid[all_faces]
    blocks = []
    for tag, axis, value in [(2, 1, 0), (3, 0, dims[0]), (4, 1, dims[1]), (5, 0, 0), (6, 2, dims[2]), (7, 2, 0)]:
        tri = all_faces[np.all(fg[:, :, axis] == value, axis=1)]
        blocks.append(("triangle", tri, np.full(len(tri), tag, dtype=np.int64)))
    blocks.append(("triangle", positive, np.full(len(positive), 50, dtype=np.int64)))
    blocks.append(("triangle", negative, np.full(len(negative), 50, dtype=np.int64)))
    blocks.append(("line", c1_boundary, np.full(len(c1_boundary), 30, dtype=np.int64)))
    blocks.append(("triangle", c1, np.full(len(c1), 10, dtype=np.int64)))
    blocks.append(("tetra", new_tets, np.ones(len(new_tets), dtype=np.int64)))
    mesh = Mesh.from_arrays(3, points, blocks)
    mesh.build_conformal_mesh()
    return mesh, {"c1": c1, "positive": positive, "negative": negative}
