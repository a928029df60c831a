"""Whole-mesh orientation data for elements whose dof transformations are not the identity: the vectorised form of
``permute_and_transform`` (R/src/basis/permute_and_transform.py:8-149).

The reference decides, per cell and per sub-entity, which matrix to apply to that sub-entity's basis functions.
Here the decision is a small integer ("case") per (cell, facet), computed for all cells at once, and the matrices
form a table indexed by the case; the CUDA kernels apply ``T[case]`` while they map the basis.

3-D faces (``_validate_face_orientation_3d``, :57-107): ``pos`` = index of the face's node tags among the
lexicographic permutations of the sorted tags -> (rotations, reflections) = 0:(0,0) 1:(0,1) 2:(2,1) 3:(1,0) 4:(2,0)
5:(1,1); a face carrying a physical tag with local index 1 or 3 is forced to ``reflections = 1`` (:102-105);
``T = reflect^reflections . rotate^rotations`` (:117-123); case id = 2 * rotations + reflections.
2-D edges (``_validate_edge_orientation_2d``, :8-19): oriented iff the cell's local edge vertices equal the edge
cell's own ``node_tags`` (order included); case id = 0 (oriented) or 1 (apply ``transformations["interval"][0]``).
"""
import numpy as np

from ..mesh.mesh import TET_FACES, TRI_EDGES, NO_MATERIAL

_FACE_CASE = np.array([[0, 0], [0, 1], [2, 1], [1, 0], [2, 0], [1, 1]])


def facet_cases(mesh, dim, selected_local):
    """``int8[n_cells, dim+1]`` case id of every facet of every selected cell."""
    if dim == 3:
        tets = mesh.vol_nodes[selected_local]
        faces = mesh.vol_faces[selected_local]
        out = np.zeros((len(tets), 4), dtype=np.int8)
        for f, fc in enumerate(TET_FACES):
            tags = tets[:, fc]
            rank = np.argsort(np.argsort(tags, axis=1), axis=1)
            pos = 2 * rank[:, 0] + (rank[:, 1] > rank[:, 2])
            rot, ref = _FACE_CASE[pos, 0].copy(), _FACE_CASE[pos, 1].copy()
            if f in (1, 3):
                tagged = mesh.material[faces[:, f]] != NO_MATERIAL
                ref[tagged] = 1
            out[:, f] = 2 * rot + ref
        return out
    if dim == 2:
        tris = mesh.face_nodes[selected_local]
        edges = mesh.face_edges[selected_local]
        out = np.zeros((len(tris), 3), dtype=np.int8)
        for e, (a, b) in enumerate(TRI_EDGES):
            own = mesh.edge_nodes[edges[:, e] - mesh.offset[1]]
            oriented = (tris[:, a] == own[:, 0]) & (tris[:, b] == own[:, 1])
            out[:, e] = np.where(oriented, 0, 1)
        return out
    raise NotImplementedError("orientation cases for dimension %d" % dim)


def edge_signs(mesh, dim, selected_local):
    """``int8[n_cells, n_edges]`` factor of every edge dof of the lowest-order Nedelec element ("N1E":
    ``transformations["interval"][0] = [-1]``): -1 for an edge the reference finds mis-oriented, +1 otherwise.
    2-D (R/src/basis/permute_and_transform.py:8-19): oriented iff the cell's local edge vertices equal the edge cell's
    own ``node_tags``, order included; 3-D (:38-54): oriented iff they equal the SORTED tags, i.e. ascend."""
    if dim == 2:
        return np.where(facet_cases(mesh, 2, selected_local) == 0, 1, -1).astype(np.int8)
    if dim == 3:
        from ..mesh.mesh import TET_EDGES

        tets = mesh.vol_nodes[selected_local]
        out = np.empty((len(tets), 6), dtype=np.int8)
        for e, (a, b) in enumerate(TET_EDGES):
            out[:, e] = np.where(tets[:, a] < tets[:, b], 1, -1)
        return out
    raise NotImplementedError("edge orientation for dimension %d" % dim)


def edge_dof_signs(mesh, dim, selected_local, ref_element):
    """``int8[n_cells, n_phi]`` factor of every dof of an edge-based H(curl) element whose edge transformation is diagonal
    (N1E_1: ``[-1]``; N2E_1: ``diag(-1, +1)``): the diagonal entry for a mis-oriented edge, +1 otherwise."""
    flip = np.diag(np.asarray(ref_element.entity_transformations()["interval"][0])).astype(np.int8)
    s = edge_signs(mesh, dim, selected_local)  # [n_cells, n_edges]
    out = np.ones((s.shape[0], s.shape[1] * len(flip)), dtype=np.int8)
    for k, f in enumerate(flip):
        out[:, k :: len(flip)] = np.where(s < 0, f, 1)
    return out


def case_matrices(ref_element):
    """``float64[n_cases, nfd, nfd]`` with ``T[0] = I`` (n_cases = 6 in 3-D, 2 in 2-D)."""
    dim = ref_element.dim
    T = ref_element.entity_transformations()
    if dim == 3:
        rotate, reflect = T["triangle"]
        n = len(rotate)
        out = np.zeros((6, n, n))
        for rot in range(3):
            for ref in range(2):
                M = np.identity(n)
                for _ in range(rot):
                    M = rotate @ M
                for _ in range(ref):
                    M = reflect @ M
                out[2 * rot + ref] = M
        return out
    if dim == 2:
        flip = T["interval"][0]
        return np.stack([np.identity(len(flip)), flip])
    raise NotImplementedError("orientation matrices for dimension %d" % dim)
