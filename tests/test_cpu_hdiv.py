"""CPU tests of the restated H(div) elements: product (exact rational) vs oracle (numerical) tables, the dual
functionals, orientation tables, and H(div)-conformity of the global basis under the reference's orientation rule."""
from collections import defaultdict

import numpy as np
import pytest

from nfempy_b200.basis import create_element, make_quadrature
from nfempy_b200.basis.orientation import case_matrices, facet_cases
from nfempy_b200.basis.reference_element import REF_VERTICES, SUB_ENTITIES
from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
from oracle import ref_basis, ref_forms, ref_hdiv
from oracle.ref_mesh import RefMesh
from oracle.ref_spaces import RefProductSpace

ELEMENTS = [("RT", 2, 1), ("RT", 2, 2), ("RT", 2, 3), ("RT", 3, 1), ("RT", 3, 2), ("RT", 3, 3), ("BDM", 2, 1), ("BDM", 3, 1),
            ("BDM", 2, 2), ("BDM", 3, 2), ("BDM", 2, 3), ("BDM", 3, 3)]


@pytest.mark.parametrize("family,dim,k", ELEMENTS)
def test_tables_match_oracle_and_layout(family, dim, k):
    el = create_element(family, dim, k)
    pts, _ = make_quadrature(dim, 2 * k + 1)
    a, b = el.tabulate(1, pts), ref_hdiv.tabulate(family, dim, k, pts)
    # the product builds its dual matrix in exact rationals; the oracle snaps its quadrature moments to the exact
    # rationals and solves in 60-digit arithmetic (oracle/ref_hdiv.py _solve_dual): both are correct to the last bits
    assert np.abs(a - b).max() <= 1e-14 * np.abs(b).max()
    assert el.num_entity_dofs == ref_basis.num_entity_dofs(family, dim, k, False)
    assert el.entity_dofs == ref_hdiv.entity_dofs(family, dim, k)
    assert not el.dof_transformations_are_identity


@pytest.mark.parametrize("family,dim,k", ELEMENTS)
def test_divergence_theorem_and_lowest_order_normal_moments(family, dim, k):
    el = create_element(family, dim, k)
    pts, w = make_quadrature(dim, 2 * k + 1)
    tab = el.tabulate(1, pts)
    lhs = (w[:, None] * sum(tab[1 + d, :, :, d] for d in range(dim))).sum(axis=0)
    verts = REF_VERTICES[dim]
    fp, fw = make_quadrature(dim - 1, 2 * k + 1)
    rhs = np.zeros(el.n_phi)
    cen = verts.mean(axis=0)
    for f, lv in enumerate(SUB_ENTITIES[dim][dim - 1]):
        v0 = verts[lv[0]]
        E = np.array([verts[lv[j]] - v0 for j in range(1, dim)])
        n = np.array([-E[0][1], E[0][0]]) if dim == 2 else np.cross(E[0], E[1])
        flux = (fw[:, None] * (el.tabulate(0, v0 + fp @ E)[0] @ n)).sum(axis=0)
        if family == "RT" and k == 1:  # biorthogonality of the single normal moment per facet
            expect = np.zeros(el.n_phi)
            expect[f] = 1.0
            assert np.abs(flux - expect).max() < 1e-14
        rhs += flux * (1.0 if n @ (v0 - cen) > 0 else -1.0)
    assert np.abs(lhs - rhs).max() <= 1e-12 * max(1.0, np.abs(lhs).max())


@pytest.mark.parametrize("family,k", [("RT", 1), ("RT", 2), ("RT", 3), ("BDM", 1), ("BDM", 2), ("BDM", 3)])
def test_orientation_tables_match_oracle(family, k):
    T3 = case_matrices(create_element(family, 3, k))
    rot, ref = ref_hdiv.entity_transformations(family, 3, k)["triangle"]
    for a in range(3):
        for b in range(2):
            assert np.array_equal(T3[2 * a + b], np.linalg.matrix_power(ref, b) @ np.linalg.matrix_power(rot, a))
    T2 = case_matrices(create_element(family, 2, k))
    assert np.array_equal(T2[1], ref_hdiv.entity_transformations(family, 2, k)["interval"][0])


def _oracle_space(mesh, family, k):
    dim = mesh.dimension
    r = RefMesh(dim, mesh.points, mesh._blocks)
    r.build_conformal_mesh()
    sp = RefProductSpace({"u": (dim, 1, family, k, r)})
    sp.build_structures({"u": [1]}, {"u": [2, 3]})
    return r, sp


def test_facet_cases_match_oracle_decisions():
    """Vectorised case ids reproduce permute_and_transform's per-face decisions (incl. the tagged-face rule)."""
    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True)
    r, sp = _oracle_space(mesh, "RT", 2)
    us = sp.discrete_spaces["u"]
    cases = facet_cases(mesh, 3, np.arange(mesh.n_volumes))
    T = case_matrices(create_element("RT", 3, 2))
    pts = np.array([[0.2, 0.3, 0.1], [0.1, 0.1, 0.6]])
    raw = ref_hdiv.tabulate("RT", 3, 2, pts)
    for i, cid in enumerate(us.element_ids[:20]):
        el = ref_forms.RefElement(us, cid)
        got = ref_hdiv.permute_and_transform(raw.copy(), el)
        want = raw.copy()
        for f in range(4):
            want[:, :, 3 * f : 3 * f + 3, :] = np.einsum("ab,dqbc->dqac", T[cases[i, f]], raw[:, :, 3 * f : 3 * f + 3, :])
        assert np.abs(got - want).max() < 1e-13


@pytest.mark.parametrize("family,dim,k", ELEMENTS)
def test_global_basis_is_hdiv_conforming(family, dim, k):
    """Normal components of every global facet function agree from both sides of every interior facet, and functions
    not attached to a facet have no normal trace on it.  Holds only for the rotation convention chosen in
    nfempy_b200/basis/transformations.py (the inverse convention fails this test)."""
    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    r, sp = _oracle_space(mesh, family, k)
    us = sp.discrete_spaces["u"]
    fq, _ = ref_basis.gauss_jacobi(dim - 1, 2 * k)
    f2c = defaultdict(list)
    for i, cid in enumerate(us.element_ids):
        for f in r.cells[cid].sub_cells_ids[dim - 1]:
            f2c[int(f)].append(i)
    jump = leak = 0.0
    for f, cells in f2c.items():
        fc = r.cells[f]
        P = r.points[fc.node_tags]
        nrm = np.cross(P[1] - P[0], P[2] - P[0]) if dim == 3 else np.array([-(P[1] - P[0])[1], (P[1] - P[0])[0], 0.0])
        nrm /= np.linalg.norm(nrm)
        fdofs = us.dof_map.maps[dim - 1][f]
        flux = []
        for i in cells:
            el = ref_forms.RefElement(us, us.element_ids[i])
            pts = ref_forms.transform_lower_to_higher(fq, fc, el.cell)
            x, jac, det, inv = el.evaluate_mapping(pts)
            tab = el.evaluate_basis(pts, jac, det, inv)
            dest = sp.destination_indexes(i)
            fl = tab[0] @ nrm
            loc = [int(np.where(dest == g)[0][0]) for g in fdofs]
            flux.append(fl[:, loc])
            rest = [j for j in range(len(dest)) if j not in loc]
            leak = max(leak, np.abs(fl[:, rest]).max())
        if len(flux) == 2:
            jump = max(jump, np.abs(flux[0] - flux[1]).max())
    scale = float(k * k * 50)
    assert jump <= 1e-12 * scale and leak <= 1e-11 * scale


@pytest.mark.parametrize("dim", [2, 3])
def test_bdm3_moment_families(dim):
    """The two families BDM_3's dual basis is built from: the degree-2 Nedelec test functions span [P1]^d + {p in
    [Ptilde_2]^d : p . x = 0} (dimension 8 / 20), and the cubic facet lattice is invariant under the facet symmetries
    (so the orientation matrices are signed permutations)."""
    from fractions import Fraction

    from nfempy_b200.basis.hdiv import _pmul, _padd, cubic_lattice, nedelec2_test_functions
    from nfempy_b200.basis.transformations import _node_permutation

    fns = nedelec2_test_functions(dim)
    assert len(fns) == (8 if dim == 2 else 20)
    n_p1 = dim * (dim + 1)
    for w in fns[n_p1:]:  # homogeneous quadratic part: x . p = 0 identically
        dot = {}
        for c in range(dim):
            unit = tuple(1 if d == c else 0 for d in range(dim))
            dot = _padd(dot, _pmul(w[c], {unit: Fraction(1)}))
        assert all(v == 0 for v in dot.values())
        assert all(sum(e) == 2 for comp in w for e, v in comp.items() if v != 0)
    # linear independence: evaluate at random points and take the rank
    rng = np.random.default_rng(1)
    pts = rng.random((40, dim))
    M = np.zeros((len(fns), 40 * dim))
    for i, w in enumerate(fns):
        for c in range(dim):
            val = np.zeros(40)
            for e, v in w[c].items():
                val += float(v) * np.prod(pts ** np.array(e), axis=1)
            M[i, c::dim] = val
    assert np.linalg.matrix_rank(M) == len(fns)
    fdim = dim - 1
    nodes = cubic_lattice(fdim)
    assert len(nodes) == (4 if fdim == 1 else 10)
    images = [[1, 0]] if fdim == 1 else [[1, 2, 0], [0, 2, 1]]
    for image in images:
        P = _node_permutation(fdim, 3, image)
        assert np.all(P.sum(axis=0) == 1) and np.all(P.sum(axis=1) == 1)
