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
