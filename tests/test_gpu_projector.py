"""GPU tests of ``L2ProjectorWeakForm``: parity with the oracle, and the reference's own projector property test
(R/tests/test_projectors/test_projectors.py:86-216) on real cells: the L2 projection of a polynomial that lies in the
space reproduces it (error <= 1e-12 at quadrature points)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
from oracle import ref_assembly, ref_forms
from tests.parity_helpers import Case, compare_csr

pytestmark = pytest.mark.gpu


def _scalar(x, y, z):
    return np.array([1.0 + 2.0 * x - y + 0.5 * z])


def _vector(x, y, z):  # [n_comp = 1, 3, n]: a linear field (in RT_2 / BDM_1)
    return np.array([[1.0 + x - 2.0 * y, 0.5 * y + z, x - z + 0.25]])


def _case(dim, fields):
    from nfempy_b200.weak_forms import L2ProjectorWeakForm

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    functions = {n: (_vector if fam in ("RT", "BDM") else _scalar) for n, (fam, k, disc) in fields.items()}

    def gpu_forms(space, a):
        wf = L2ProjectorWeakForm(space)
        wf.functions = functions
        first = next(iter(space.discrete_spaces.values()))
        return {"proj": ["form", range(len(first.elements)), wf, a]}

    def ref_forms_fn(space):
        first = next(iter(space.discrete_spaces.values()))
        return [("form", ref_forms.l2_projector, functions, len(first.element_ids))]

    return Case(
        mesh, lambda g: {n: (dim, 1, fam, k, g) for n, (fam, k, disc) in fields.items()}, {n: [1] for n in fields},
        {}, gpu_forms, ref_forms_fn, disc={n: disc for n, (fam, k, disc) in fields.items()},
    )


@pytest.mark.parametrize("dim", [2, 3])
def test_projector_matches_oracle_mixed_fields(dim):
    case = _case(dim, {"u": ("RT", 2, False), "p": ("Lagrange", 1, True), "q": ("Lagrange", 2, False)})
    compare_csr(case.run_gpu(), case.run_oracle(), 1e-12)


@pytest.mark.parametrize("dim,fam,k,disc", [
    (2, "Lagrange", 1, False), (3, "Lagrange", 2, False), (3, "Lagrange", 1, True), (2, "RT", 2, False),
    (3, "RT", 2, False), (3, "RT", 3, False), (3, "BDM", 1, False), (2, "BDM", 1, False),
])
def test_projection_of_polynomial_in_the_space_is_exact(dim, fam, k, disc):
    """The reference's projector test, on real cells: project, then compare u_h with u at the quadrature points."""
    case = _case(dim, {"u": (fam, k, disc)})
    indptr, indices, data, rhs = case.run_gpu()
    alpha = spl.spsolve(sp.csr_matrix((data, indices, indptr)).tocsc(), -rhs)
    space = case.ref_space()
    us = space.discrete_spaces["u"]
    points, _ = space.quadrature[dim]
    worst = 0.0
    for i, cid in enumerate(us.element_ids):
        el = ref_forms.RefElement(us, cid)
        x, jac, det, inv = el.evaluate_mapping(points)
        tab = el.evaluate_basis(points, jac, det, inv)
        dest = space.destination_indexes(i)
        if fam == "Lagrange":
            uh = tab[0, :, :, 0] @ alpha[dest]
            exact = _scalar(x[:, 0], x[:, 1], x[:, 2])[0]
        else:
            uh = np.einsum("qjc,j->qc", tab[0, :, :, 0:dim], alpha[dest])
            exact = _vector(x[:, 0], x[:, 1], x[:, 2])[0, 0:dim, :].T
        worst = max(worst, np.abs(uh - exact).max())
    assert worst <= 1e-12


def _nedelec_field(dim):
    """A field of the lowest-order Nedelec space, ``c + d x X`` (2-D: ``(a - w y, b + w x, 0)``), as [1, 3, n]."""
    if dim == 2:
        return lambda x, y, z: np.array([[0.7 - 1.3 * y, -0.4 + 1.3 * x, 0.0 * x]])
    c, d = np.array([0.3, -1.1, 0.8]), np.array([0.5, 0.25, -0.75])
    return lambda x, y, z: np.array([[c[0] + d[1] * z - d[2] * y, c[1] + d[2] * x - d[0] * z, c[2] + d[0] * y - d[1] * x]])


@pytest.mark.parametrize("dim", [2, 3])
def test_hcurl_projector_matches_oracle(dim):
    """N1E through DiscreteSpace + L2ProjectorWeakForm (covariant map, edge-orientation signs, csrc/nfa_hcurl.cu) against
    the oracle's physical-space Whitney functions: CSR pattern bit-exact, values and right-hand side to 1e-12."""
    from nfempy_b200.weak_forms import L2ProjectorWeakForm
    from oracle import ref_hcurl

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    functions = {"e": lambda x, y, z: np.array([[np.sin(x) + y, x * y - z, 1.0 + z * x]])}

    def gpu_forms(space, a):
        wf = L2ProjectorWeakForm(space)
        wf.functions = functions
        return {"proj": ["form", range(len(space.discrete_spaces["e"].elements)), wf, a]}

    def ref_forms_fn(space):
        return [("form", ref_hcurl.l2_projector, functions, len(space.discrete_spaces["e"].element_ids))]

    case = Case(mesh, lambda g: {"e": (dim, 1, "N1E", 1, g)}, {"e": [1]}, {}, gpu_forms, ref_forms_fn)
    compare_csr(case.run_gpu(), case.run_oracle(), 1e-12)


@pytest.mark.parametrize("dim", [2, 3])
def test_hcurl_projection_of_a_field_in_the_space_is_exact(dim):
    """The reference's projector property test for the N1E family (R/tests/test_projectors/test_projectors.py:187-216):
    ``l2_projector`` of a function of the space, then ``l2_error <= 1e-12`` -- both through the CUDA path."""
    from nfempy_b200.postprocess import l2_error, l2_projector
    from nfempy_b200.spaces import ProductSpace

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    space = ProductSpace({"e": (dim, 1, "N1E", 1, mesh)})
    space.build_structures({"e": [1]}, {})
    f = {"e": _nedelec_field(dim)}
    alpha = l2_projector(space, f)
    err = l2_error(dim, space, f, alpha)
    assert len(err) == 1 and err[0] <= 1e-12, err
    # and the coefficients are the tangential moments: a wrong orientation sign would show here even if the error norm
    # were computed with the same mistake -- check one edge against the line integral of f along it
    s = space.discrete_spaces["e"]
    edge = s.mesh_topology.entity_ids[1][0]
    a, b = mesh.edge_nodes[edge - mesh.offset[1]]
    xa, xb = mesh.points[a], mesh.points[b]
    ts = (np.polynomial.legendre.leggauss(4)[0] + 1.0) / 2.0
    ws = np.polynomial.legendre.leggauss(4)[1] / 2.0
    xs = xa[None, :] + ts[:, None] * (xb - xa)[None, :]
    moment = float(np.sum(ws * np.einsum("dq,d->q", f["e"](xs[:, 0], xs[:, 1], xs[:, 2])[0], xb - xa)))
    dof = int(s.dof_map.entity_block(1, np.array([edge]))[0, 0])
    assert abs(abs(alpha[dof]) - abs(moment)) <= 1e-12 * max(1.0, abs(moment))


@pytest.mark.parametrize("dim", [2, 3])
def test_n2e_projector_matches_oracle(dim):
    """N2E of degree 1 (two tangential moments per edge, edge transformation diag(-1, +1)) through DiscreteSpace +
    L2ProjectorWeakForm + csrc/nfa_hcurl.cu against the oracle's physical-space functions."""
    from nfempy_b200.weak_forms import L2ProjectorWeakForm
    from oracle import ref_hcurl

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    functions = {"e": lambda x, y, z: np.array([[np.sin(x) + y, x * y - z, 1.0 + z * x]])}

    def gpu_forms(space, a):
        wf = L2ProjectorWeakForm(space)
        wf.functions = functions
        return {"proj": ["form", range(len(space.discrete_spaces["e"].elements)), wf, a]}

    def ref_forms_fn(space):
        return [("form", ref_hcurl.l2_projector_n2e, functions, len(space.discrete_spaces["e"].element_ids))]

    case = Case(mesh, lambda g: {"e": (dim, 1, "N2E", 1, g)}, {"e": [1]}, {}, gpu_forms, ref_forms_fn)
    compare_csr(case.run_gpu(), case.run_oracle(), 1e-12)


@pytest.mark.parametrize("dim", [2, 3])
def test_n2e_projection_of_a_linear_field_is_exact(dim):
    """The reference's projector property test for the N2E family at its first order
    (R/tests/test_projectors/test_projectors.py:187-216, ``v_functions[0] = (y, -x, -z)``), and a general affine field:
    ``l2_projector`` then ``l2_error <= 1e-12`` through the CUDA path; the even dof of an edge is the tangential moment of
    the field along the globally oriented edge."""
    from nfempy_b200.postprocess import l2_error, l2_projector
    from nfempy_b200.spaces import ProductSpace

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    space = ProductSpace({"e": (dim, 1, "N2E", 1, mesh)})
    space.build_structures({"e": [1]}, {})
    planar = 0.0 if dim == 2 else 1.0  # a 2-D space only represents the in-plane components
    fields = [
        lambda x, y, z: np.array([[y, -x, -z * planar]]),
        lambda x, y, z: np.array([[1.0 + 2.0 * x - y + 0.5 * z * planar, 0.3 - x + 4.0 * y, planar * (2.0 - 3.0 * z + x)]]),
    ]
    for fn in fields:
        f = {"e": fn}
        alpha = l2_projector(space, f)
        err = l2_error(dim, space, f, alpha)
        assert len(err) == 1 and err[0] <= 1e-12, err
        s = space.discrete_spaces["e"]
        edge = s.mesh_topology.entity_ids[1][3]
        a, b = mesh.edge_nodes[edge - mesh.offset[1]]
        xa, xb = mesh.points[a], mesh.points[b]
        ts, ws = np.polynomial.legendre.leggauss(4)
        ts, ws = (ts + 1.0) / 2.0, ws / 2.0
        xs = xa[None, :] + ts[:, None] * (xb - xa)[None, :]
        moment = float(np.sum(ws * np.einsum("dq,d->q", fn(xs[:, 0], xs[:, 1], xs[:, 2])[0], xb - xa)))
        dof = int(s.dof_map.entity_block(1, np.array([edge]))[0, 0])
        assert abs(abs(alpha[dof]) - abs(moment)) <= 1e-12 * max(1.0, abs(moment))
