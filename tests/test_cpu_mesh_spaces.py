"""CPU tests: array-based mesh numbering, DoF numbering and destination tables vs the cell-by-cell oracle, and
vs the committed numbering fixture of the reference's own gmsh mesh."""
import numpy as np
import pytest

from nfempy_b200.mesh import Mesh, box_tet_mesh, square_tri_mesh
from nfempy_b200.spaces import ProductSpace
from oracle.ref_mesh import RefMesh
from oracle.ref_spaces import RefProductSpace
from tests.golden_cases import GOLDEN, load_mesh_blocks


def _same_cells(mesh):
    ref = RefMesh(mesh.dimension, mesh.points, mesh._blocks)
    ref.build_conformal_mesh()
    assert len(ref.cells) == mesh.n_cells
    for c in ref.cells:
        if c is None:
            continue
        p = mesh.cells[c.id]
        assert p.dimension == c.dimension and p.material_id == c.material_id
        assert np.array_equal(p.node_tags, c.node_tags)
        for d in range(c.dimension + 1):
            assert np.array_equal(p.sub_cells_ids[d], c.sub_cells_ids[d])
    return ref


@pytest.mark.parametrize("shuffle", [False, True])
def test_box_mesh_numbering_matches_oracle(shuffle):
    _same_cells(box_tet_mesh(3, shuffle_local_nodes=shuffle))


def test_square_mesh_numbering_matches_oracle():
    _same_cells(square_tri_mesh(5, jitter=0.2))


def test_reference_gmsh_fixture_numbering():
    points, blocks = load_mesh_blocks("mesh_ex3_3d_l0.npz")
    mesh = Mesh.from_arrays(3, points, blocks)
    mesh.build_conformal_mesh()
    z = np.load(GOLDEN + "/numbering_ex3_3d_l0.npz")
    assert mesh.n_cells == len(z["dims"])
    assert np.array_equal(mesh.cell_dimension(np.arange(mesh.n_cells)), z["dims"])
    assert np.array_equal(mesh.material, z["material"])
    assert np.array_equal(mesh.vol_edges, z["vol_edges"]) and np.array_equal(mesh.vol_faces, z["vol_faces"])
    for dim in (1, 2, 3):
        ids = mesh.ids_of_dimension(dim)
        assert np.array_equal(mesh.node_tags_of(dim), z["node_tags"][ids, : dim + 1])
    _same_cells(mesh)


def test_mixed_block_order_materialisation():
    """Tets listed BEFORE boundary triangles: sub-entities keep the tet's local node order (mesh.py:180-219)."""
    m = box_tet_mesh(2)
    blocks = [b for b in m._blocks if b[0] == "tetra"] + [b for b in m._blocks if b[0] != "tetra"]
    m2 = Mesh.from_arrays(3, m.points, blocks)
    m2.build_conformal_mesh()
    _same_cells(m2)


def _check_space(mesh, data_fn, tags, btags, disc=None):
    ref = RefMesh(mesh.dimension, mesh.points, mesh._blocks)
    ref.build_conformal_mesh()
    ps = ProductSpace(data_fn(mesh))
    rs = RefProductSpace(data_fn(ref))
    for s in (ps, rs):
        s.make_subspaces_discontinuous(disc or {})
        s.build_structures(tags, btags)
    assert ps.n_dof == rs.n_dof
    first = ps.names[0]
    n_el = len(ps.discrete_spaces[first].elements)
    assert n_el == len(rs.discrete_spaces[first].element_ids)
    for i in range(n_el):
        assert np.array_equal(ps.destination_indexes(i), rs.destination_indexes(i))
    n_bc = len(ps.discrete_spaces[first].bc_elements)
    assert n_bc == len(rs.discrete_spaces[first].bc_element_ids)
    for i in range(n_bc):
        assert np.array_equal(ps.bc_destination_indexes(i), rs.bc_destination_indexes(i))
    return ps


def test_dof_numbering_vector_p2_3d():
    mesh = box_tet_mesh(3, shuffle_local_nodes=True)
    ps = _check_space(mesh, lambda g: {"u": (3, 3, "Lagrange", 2, g)}, {"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
    assert ps.destination_table().shape[1] == 30


def test_dof_numbering_two_fields_with_discontinuous():
    mesh = box_tet_mesh(2)
    data = lambda g: {"u": (3, 3, "Lagrange", 2, g), "p": (3, 1, "Lagrange", 1, g)}
    _check_space(mesh, data, {"u": [1], "p": [1]}, {"u": [2, 3]}, {"p": True})


@pytest.mark.parametrize("k", [1, 2])
def test_dof_numbering_2d(k):
    mesh = square_tri_mesh(4, jitter=0.1)
    _check_space(mesh, lambda g: {"p": (2, 1, "Lagrange", k, g)}, {"p": [1]}, {"p": [2, 3, 4, 5]})


def test_dof_shift_and_field_access():
    mesh = box_tet_mesh(2)
    ps = ProductSpace({"u": (3, 3, "Lagrange", 1, mesh)})
    ps.build_structures({"u": [1]}, {})
    base = ps.destination_indexes(3).copy()
    ps.dof_shift = 100
    assert np.array_equal(ps.destination_indexes(3), base + 100)
    assert np.array_equal(ps.destination_indexes(3, "u"), base + 100)
    with pytest.raises(ValueError):
        ps.dof_shift = -1
    with pytest.raises(TypeError):
        ps.dof_shift = 1.5


def test_empty_physical_tag_selects_no_cells():
    """The reference's projector test passes tags that match nothing; the space is then empty."""
    mesh = box_tet_mesh(2)
    ps = ProductSpace({"u": (3, 1, "Lagrange", 1, mesh)})
    ps.build_structures({"u": []}, {})
    assert ps.n_dof == 0 and len(ps.discrete_spaces["u"].elements) == 0
    assert ps.destination_table().shape[0] == 0


@pytest.mark.parametrize("dim", [2, 3])
def test_continuous_cubic_lagrange_is_conforming_and_matches_the_oracle(dim):
    """P_3 is the first Lagrange degree with non-identity dof transformations (two interior nodes per edge).  The product
    lists the two dofs of a mis-oriented edge in exchanged order and keeps one reference table; the oracle, like the
    reference, exchanges the two basis functions.  (a) With the product's tables the global functions are single-valued
    along every shared edge; (b) cell by cell, the function the product attaches to a global dof is the function the
    oracle attaches to it."""
    from nfempy_b200.basis.reference_element import REF_VERTICES, SUB_ENTITIES
    from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
    from oracle.ref_forms import RefElement
    from oracle.ref_mesh import RefMesh
    from oracle.ref_spaces import RefProductSpace

    mesh = square_tri_mesh(3, jitter=0.2) if dim == 2 else box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True)
    space = ProductSpace({"f": (dim, 1, "Lagrange", 3, mesh)})
    space.build_structures({"f": [1]}, {})
    fs = space.discrete_spaces["f"]
    dest = np.asarray(space.destination_table())
    raw = np.asarray(fs.dof_map._dest_cache())
    assert (dest != raw).any() and np.array_equal(np.sort(dest, axis=1), np.sort(raw, axis=1))
    cn, ref, verts = fs.cell_nodes(), fs.ref_element, REF_VERTICES[dim]
    seen, shared = {}, 0
    for i in range(len(cn)):
        for a, b in SUB_ENTITIES[dim][1]:
            ga, gb = int(cn[i][a]), int(cn[i][b])
            lo, hi = (a, b) if ga < gb else (b, a)
            for t in (0.3, 0.6):
                vals = ref.tabulate(0, (verts[lo] + t * (verts[hi] - verts[lo]))[None, :])[0, 0, :, 0]
                for j, d in enumerate(dest[i]):
                    key = (int(d), min(ga, gb), max(ga, gb), t)
                    if key in seen:
                        shared += 1
                        assert abs(seen[key] - vals[j]) < 1e-12
                    else:
                        seen[key] = vals[j]
    assert shared > 50
    r = RefMesh(dim, mesh.points, mesh._blocks)
    r.build_conformal_mesh()
    rs = RefProductSpace({"f": (dim, 1, "Lagrange", 3, r)})
    rs.build_structures({"f": [1]}, {})
    rf = rs.discrete_spaces["f"]
    pts, _ = space.quadrature[dim]
    tab_p = ref.tabulate(0, pts)[0, :, :, 0]
    for i in range(len(cn)):
        el = RefElement(rf, rf.element_ids[i])
        x, jac, det_jac, inv_jac = el.evaluate_mapping(pts)
        tab_o = el.evaluate_basis(pts, jac, det_jac, inv_jac)[0, :, :, 0]
        d_o = list(rs.destination_indexes(i))
        assert sorted(d_o) == sorted(dest[i].tolist())
        for j, d in enumerate(dest[i]):
            assert np.abs(tab_p[:, j] - tab_o[:, d_o.index(int(d))]).max() < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
def test_cubic_lagrange_boundary_elements_see_the_dofs_where_the_cells_do(dim):
    """Every global dof of a continuous P_3 field sits at one physical node, whether it is reached through a cell's
    destination row and the cell's reference nodes or through a boundary element's destination row and the facet's
    reference nodes (Dirichlet penalty forms integrate the facet's own nodal basis against those dofs)."""
    from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh

    mesh = square_tri_mesh(3, jitter=0.2) if dim == 2 else box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True)
    space = ProductSpace({"f": (dim, 2, "Lagrange", 3, mesh)})
    space.build_structures({"f": [1]}, {"f": [2, 3, 4, 5] if dim == 2 else [2, 3, 4, 5, 6, 7]})
    fs = space.discrete_spaces["f"]

    def positions(ref_nodes, cell_nodes):
        lam = np.concatenate([1.0 - ref_nodes.sum(axis=1, keepdims=True), ref_nodes], axis=1)
        return np.repeat(np.einsum("jv,nvc->njc", lam, mesh.points[cell_nodes]), 2, axis=1)  # two components per node

    where = {}
    dest = np.asarray(space.destination_table())
    for row, X in zip(dest, positions(fs.ref_element.nodes, fs.cell_nodes())):
        for d, x in zip(row, X):
            assert np.abs(where.setdefault(int(d), x) - x).max() < 1e-12
    bdest = np.asarray(space.bc_destination_table())
    assert len(bdest) > 10
    for row, X in zip(bdest, positions(fs.bc_ref_element.nodes, fs.bc_cell_nodes())):
        for d, x in zip(row, X):
            assert np.abs(where[int(d)] - x).max() < 1e-12
