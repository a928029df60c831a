"""Primal forms on several Lagrange fields (LCEPrimalWeakForm + BC, LETaylorHoodWeakForm + BC,
LEPrimalStressConstraintWeakForm) through the generic multi-field kernel, against the cell-by-cell oracle
(oracle/ref_h1_multi.py): bit-exact pattern, values and residual within 1e-12."""
import numpy as np
import pytest

from tests.parity_helpers import compare_csr, lce_primal_case, stress_constraint_case, taylor_hood_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.mark.parametrize("dim,k,kt,varying,alpha", [(3, 2, 2, False, False), (3, 2, 1, True, True), (3, 1, 1, True, False),
                                                     (2, 2, 2, True, True), (2, 2, 1, False, False)])
def test_lce_primal(dim, k, kt, varying, alpha):
    case = lce_primal_case(dim=dim, k=k, kt=kt, n=2 if dim == 3 else 3, varying=varying, alpha=alpha)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


@pytest.mark.parametrize("varying,alpha", [(False, False), (True, True)])
def test_taylor_hood_2d(varying, alpha):
    case = taylor_hood_case(n=3, varying=varying, alpha=alpha)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_taylor_hood_3d_is_rejected_like_the_reference():
    from nfempy_b200.mesh import box_tet_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LETaylorHoodWeakForm

    mesh = box_tet_mesh(1)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh), "p": (3, 1, "Lagrange", 1, mesh)})
    space.build_structures({"u": [1], "p": [1]}, {})
    wf = LETaylorHoodWeakForm(space)
    wf.functions = {"rhs": lambda x, y, z: np.zeros((3,) + np.shape(x)), "lambda": lambda x, y, z: 1.0 + 0 * x,
                    "mu": lambda x, y, z: 1.0 + 0 * x}
    with pytest.raises(NotImplementedError):  # the reference's 3-D branch starts with `assert False`
        wf.evaluate_form(0, np.zeros(34))


@pytest.mark.parametrize("dim,alpha", [(3, False), (3, True), (2, True)])
def test_stress_constraint(dim, alpha):
    case = stress_constraint_case(dim=dim, n=2 if dim == 3 else 3, alpha=alpha)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_evaluate_form_matches_the_oracle_cell_by_cell():
    """The reference's per-cell API: ``evaluate_form(i, alpha_l) -> (r_el, j_el)``."""
    from nfempy_b200.weak_forms import LCEPrimalWeakForm
    from oracle import ref_h1_multi

    case = lce_primal_case(dim=3, k=2, kt=1, n=1, varying=True)
    space, rspace = case.space(), case.ref_space()
    wf = case.gpu_forms(space, None)["volume"][2]
    assert isinstance(wf, LCEPrimalWeakForm)
    functions = case.ref_forms_fn(rspace)[0][2]
    rng = np.random.default_rng(0)
    for i in (0, 3):
        alpha_l = rng.standard_normal(wf.n_loc())
        r, j = wf.evaluate_form(i, alpha_l)
        r_ref, j_ref = ref_h1_multi.lce_primal(rspace, functions, i, alpha_l)
        assert np.abs(j - j_ref).max() <= RTOL * np.abs(j_ref).max()
        assert np.abs(r - r_ref).max() <= RTOL * np.abs(r_ref).max()
