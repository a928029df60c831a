"""H(curl) reference element N1E_1 (nfempy_b200/basis/hcurl.py): the Whitney functions, their tangential moments, and
tangential continuity of the covariantly mapped functions across physical cells once the reference's edge-orientation
rule (ascending global vertex ids, R/src/basis/permute_and_transform.py:38-54) is applied with
``entity_transformations()["interval"][0] = [[-1]]``."""
import numpy as np
import pytest

from nfempy_b200.basis.hcurl import covariant_piola
from nfempy_b200.basis.reference_element import REF_VERTICES, SUB_ENTITIES, create_element
from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh


@pytest.mark.parametrize("dim", [2, 3])
def test_basis_is_the_whitney_family_with_unit_tangential_moments(dim):
    el = create_element("N1E", dim, 1)
    edges = SUB_ENTITIES[dim][1]
    assert el.n_phi == len(edges) and el.value_size == dim and not el.dof_transformations_are_identity
    rng = np.random.default_rng(3)
    pts = rng.random((20, dim)) / dim
    tab = el.tabulate(1, pts)[0]  # [nq, n_phi, dim]
    lam = np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)
    glam = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)
    for i, (a, b) in enumerate(edges):
        whitney = lam[:, [a]] * glam[b][None, :] - lam[:, [b]] * glam[a][None, :]
        assert np.abs(tab[:, i, :] - whitney).max() < 1e-14
    verts = REF_VERTICES[dim]
    s = (np.arange(8) + 0.5) / 8.0  # midpoint rule is exact for the (affine) tangential component
    for j, (a, b) in enumerate(edges):
        x = verts[a][None, :] + s[:, None] * (verts[b] - verts[a])[None, :]
        moments = np.einsum("qic,c->i", el.tabulate(0, x)[0], verts[b] - verts[a]) / len(s)
        assert np.abs(moments - np.eye(len(edges))[j]).max() < 1e-14


@pytest.mark.parametrize("dim", [2, 3])
def test_tangential_continuity_across_cells(dim):
    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    el = create_element("N1E", dim, 1)
    flip = el.entity_transformations()["interval"][0][0, 0]
    cells = mesh.vol_nodes if dim == 3 else mesh.face_nodes
    edges = SUB_ENTITIES[dim][1]
    verts = REF_VERTICES[dim]
    seen = {}
    checked = 0
    for c, nodes in enumerate(cells):
        X = mesh.points[nodes]
        J = (X[1:] - X[0]).T  # [3, dim]
        for i, (a, b) in enumerate(edges):
            ga, gb = int(nodes[a]), int(nodes[b])
            sign = 1.0 if ga < gb else flip  # the reference transforms the functions of mis-oriented edges
            xi = 0.3 * verts[a] + 0.7 * verts[b] if ga < gb else 0.7 * verts[a] + 0.3 * verts[b]  # same physical point
            phi = covariant_piola(el.tabulate(0, xi[None, :])[0][0], J)[i] * sign  # [3]
            lo, hi = (ga, gb) if ga < gb else (gb, ga)
            tangent = mesh.points[hi] - mesh.points[lo]
            t_comp = float(phi @ tangent)
            key = (lo, hi)
            if key in seen:
                assert abs(t_comp - seen[key]) < 1e-12, (key, t_comp, seen[key])
                checked += 1
            else:
                seen[key] = t_comp
            assert abs(t_comp - 1.0) < 1e-12  # unit tangential moment along the globally oriented edge
    assert checked > 10


@pytest.mark.parametrize("dim", [2, 3])
def test_mapped_oriented_basis_equals_the_physical_whitney_functions(dim):
    """What the CUDA kernels of csrc/nfa_hcurl.cu evaluate -- reference tables, covariant map, edge signs from
    basis/orientation.edge_signs -- against the oracle's Whitney functions written directly in physical coordinates, on
    every cell of a jittered mesh with shuffled local vertices; and the N1E destination tables of product and oracle."""
    from nfempy_b200.basis.hcurl import covariant_piola
    from nfempy_b200.basis.orientation import edge_signs
    from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
    from nfempy_b200.spaces import ProductSpace
    from oracle import ref_hcurl
    from oracle.ref_mesh import RefMesh
    from oracle.ref_spaces import RefProductSpace

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    space = ProductSpace({"e": (dim, 1, "N1E", 1, mesh)})
    space.build_structures({"e": [1]}, {})
    s = space.discrete_spaces["e"]
    r = RefMesh(dim, mesh.points, mesh._blocks)
    r.build_conformal_mesh()
    rs = RefProductSpace({"e": (dim, 1, "N1E", 1, r)})
    rs.build_structures({"e": [1]}, {})
    dest = np.asarray(space.destination_table())
    assert np.array_equal(dest, np.array([rs.destination_indexes(i) for i in range(len(dest))]))
    pts, _ = space.quadrature[dim]
    sg = edge_signs(mesh, dim, s.mesh_topology.selected_local)
    tab = s.ref_element.tabulate(0, pts)[0]
    cn = s.cell_nodes()
    assert (sg == -1).any() and (sg == 1).any()
    for i in range(len(cn)):
        X = mesh.points[cn[i]]
        phi = covariant_piola(tab, (X[1:] - X[0]).T) * sg[i][None, :, None]
        _, _, ref = ref_hcurl.whitney(rs.discrete_spaces["e"], rs.discrete_spaces["e"].element_ids[i], pts)
        assert np.abs(phi - ref[:, :, : phi.shape[2]]).max() <= 1e-12


@pytest.mark.parametrize("dim", [2, 3])
def test_n2e_reference_element_is_dual_to_the_legendre_edge_moments(dim):
    """N2E of degree 1: the full [P_1]^d, two dofs per edge, dual to the tangential moments against 1 and sqrt(3)(2s - 1);
    the edge transformation is diag(-1, +1)."""
    from nfempy_b200.basis.reference_element import REF_VERTICES, SUB_ENTITIES

    el = create_element("N2E", dim, 1)
    n_edges = len(SUB_ENTITIES[dim][1])
    assert el.n_phi == 2 * n_edges == dim * (dim + 1) and el.value_size == dim
    assert el.entity_dofs[1] == [[2 * e, 2 * e + 1] for e in range(n_edges)]
    assert all(len(d) == 0 for k in (0, 2, 3)[: dim] for d in el.entity_dofs[k] if k != 1)
    s, w = np.polynomial.legendre.leggauss(4)
    s, w = 0.5 * (s + 1.0), 0.5 * w
    D = np.zeros((el.n_phi, el.n_phi))
    for e, (a, b) in enumerate(SUB_ENTITIES[dim][1]):
        va, vb = REF_VERTICES[dim][a], REF_VERTICES[dim][b]
        tab = el.tabulate(0, va[None, :] + s[:, None] * (vb - va)[None, :])[0]
        for k, q in enumerate((np.ones_like(s), np.sqrt(3.0) * (2.0 * s - 1.0))):
            D[2 * e + k] = np.einsum("q,qjc,c->j", w * q, tab, vb - va)
    assert np.abs(D - np.eye(el.n_phi)).max() <= 1e-14
    assert np.array_equal(el.entity_transformations()["interval"][0], np.diag([-1.0, 1.0]))


@pytest.mark.parametrize("dim", [2, 3])
def test_n2e_mapped_oriented_basis_equals_the_physical_functions(dim):
    """Reference tables + covariant map + per-dof edge signs (what csrc/nfa_hcurl.cu evaluates) against the oracle's
    second-kind functions written in physical coordinates, on every cell of a jittered, locally shuffled mesh; the
    destination tables of product and oracle (two dofs per edge, first-appearance edge order)."""
    from nfempy_b200.basis.hcurl import covariant_piola
    from nfempy_b200.basis.orientation import edge_dof_signs
    from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
    from nfempy_b200.spaces import ProductSpace
    from oracle import ref_hcurl
    from oracle.ref_mesh import RefMesh
    from oracle.ref_spaces import RefProductSpace

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    space = ProductSpace({"e": (dim, 1, "N2E", 1, mesh)})
    space.build_structures({"e": [1]}, {})
    s = space.discrete_spaces["e"]
    r = RefMesh(dim, mesh.points, mesh._blocks)
    r.build_conformal_mesh()
    rs = RefProductSpace({"e": (dim, 1, "N2E", 1, r)})
    rs.build_structures({"e": [1]}, {})
    dest = np.asarray(space.destination_table())
    assert dest.shape[1] == dim * (dim + 1)
    assert np.array_equal(dest, np.array([rs.destination_indexes(i) for i in range(len(dest))]))
    pts, _ = space.quadrature[dim]
    sg = edge_dof_signs(mesh, dim, s.mesh_topology.selected_local, s.ref_element)
    assert (sg[:, 0::2] == -1).any() and (sg[:, 1::2] == 1).all()
    tab = s.ref_element.tabulate(0, pts)[0]
    cn = s.cell_nodes()
    for i in range(len(cn)):
        X = mesh.points[cn[i]]
        phi = covariant_piola(tab, (X[1:] - X[0]).T) * sg[i][None, :, None]
        _, _, ref = ref_hcurl.nedelec2(rs.discrete_spaces["e"], rs.discrete_spaces["e"].element_ids[i], pts)
        assert np.abs(phi - ref[:, :, : phi.shape[2]]).max() <= 1e-12


def test_n2e_tangential_traces_are_single_valued_on_shared_edges():
    """Conformity of the second-kind element on a tetrahedral mesh: after the covariant map and the diag(-1, +1) edge
    transformation, the two functions of an edge have the same tangential trace -- as a function of the position along
    the GLOBALLY oriented edge -- in every cell that shares the edge, their moments against 1 and sqrt(3)(2s - 1) are the
    identity, and the functions of all other edges have no tangential trace there."""
    dim = 3
    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True)
    el = create_element("N2E", dim, 1)
    flip = np.diag(el.entity_transformations()["interval"][0])
    edges = SUB_ENTITIES[dim][1]
    verts = REF_VERTICES[dim]
    s, w = np.polynomial.legendre.leggauss(3)
    s, w = 0.5 * (s + 1.0), 0.5 * w
    q = np.stack([np.ones_like(s), np.sqrt(3.0) * (2.0 * s - 1.0)])
    seen, checked = {}, 0
    for nodes in mesh.vol_nodes:
        X = mesh.points[nodes]
        J = (X[1:] - X[0]).T
        for i, (a, b) in enumerate(edges):
            ga, gb = int(nodes[a]), int(nodes[b])
            oriented = ga < gb
            lo, hi = (a, b) if oriented else (b, a)  # local vertices in the global direction of the edge
            xi = verts[lo][None, :] + s[:, None] * (verts[hi] - verts[lo])[None, :]
            phi = covariant_piola(el.tabulate(0, xi)[0], J)  # [nq, n_phi, 3]
            sign = np.ones(el.n_phi)
            for e, (ea, eb) in enumerate(edges):
                if not int(nodes[ea]) < int(nodes[eb]):
                    sign[2 * e : 2 * e + 2] = flip
            tangent = mesh.points[max(ga, gb)] - mesh.points[min(ga, gb)]
            trace = np.einsum("qjc,c->qj", phi, tangent) * sign[None, :]  # [nq, n_phi] along the global direction
            moments = np.einsum("kq,q,qj->kj", q, w, trace)  # [2, n_phi]
            want = np.zeros((2, el.n_phi))
            want[0, 2 * i], want[1, 2 * i + 1] = 1.0, 1.0
            assert np.abs(moments - want).max() < 1e-12
            own = trace[:, 2 * i : 2 * i + 2]
            key = (min(ga, gb), max(ga, gb))
            if key in seen:
                assert np.abs(own - seen[key]).max() < 1e-12
                checked += 1
            else:
                seen[key] = own
    assert checked > 50
