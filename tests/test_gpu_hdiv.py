"""GPU parity: mixed Poisson on RT_k x P_{k-1}^disc (config 3 style) and BDM_1 x P_0 vs the oracle."""
import numpy as np
import pytest

from tests.parity_helpers import compare_csr, mixed_poisson_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.mark.parametrize("k", [1, 2, 3])
def test_mixed_poisson_rt_3d(k):
    case = mixed_poisson_case(dim=3, k=k, n=2 if k < 3 else 1)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_mixed_poisson_rt_2d(k):
    case = mixed_poisson_case(dim=2, k=k, n=3)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


@pytest.mark.parametrize("dim,k", [(2, 1), (3, 1), (2, 2), (3, 2), (2, 3), (3, 3)])
def test_mixed_poisson_bdm(dim, k):
    """BDM_3 (k = 3): cubic facet families (4 / 10 dofs per facet with their orientation matrices) and interior moments
    against the degree-2 Nedelec space; its oracle inverts a 20 x 20 / 60 x 60 dual matrix numerically and is the less accurate
    side (see tests/test_cpu_hdiv.py), hence the wider tolerance for that case."""
    case = mixed_poisson_case(dim=dim, k=k, n=2 if k == 1 else (1 if dim == 3 else 2), family="BDM", varying=(k == 3))
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_mixed_poisson_varying_kappa_no_bc():
    case = mixed_poisson_case(dim=3, k=2, n=2, with_bc=False, varying=True)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_mixed_poisson_solves_linear_pressure_exactly():
    """End-to-end property (the reference's projector-style check): with f = 0 and p_D = x + 2y - z the discrete
    pressure dofs of RT_2 x P_1 reproduce the linear field; needs a conforming, correctly oriented RT basis."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.weak_forms import LaplaceDualWeakForm, LaplaceDualWeakFormBCDirichlet

    case = mixed_poisson_case(dim=3, k=2, n=2)
    space = case.space()
    wf, bf = LaplaceDualWeakForm(space), LaplaceDualWeakFormBCDirichlet(space)
    wf.functions = {"rhs": lambda x, y, z: np.array([0.0 * x]), "kappa": lambda x, y, z: 2.0}
    bf.functions = {"p": lambda x, y, z: np.array([x + 2.0 * y - z])}
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    us = space.discrete_spaces["u"]
    asm.form_to_input_list = {
        "volume": ["form", range(len(us.elements)), wf, None],
        "bc": ["bc_form", range(len(us.bc_elements)), bf, None],
    }
    indptr, indices, data, rhs = asm.assemble()
    A = sp.csr_matrix((data, indices, indptr)).tocsc()
    sol = spl.spsolve(A, -rhs)
    p = sol[us.n_dof :]
    assert abs(p.min() + 1.0) < 1e-10 and abs(p.max() - 3.0) < 1e-10  # vertex values of x + 2y - z on the cube
    assert np.abs(A @ sol + rhs).max() < 1e-11


@pytest.mark.parametrize("dim,k", [(3, 1), (2, 1), (2, 2), (3, 2)])
def test_sundus_time_dependent_forms(dim, k):
    """Config 5: RT_k x RT_k x P x P implicit-Euler step with the nonlinear reaction, nonzero iterate and previous
    time level, time-dependent loads and boundary data."""
    from tests.parity_helpers import SundusCase

    case = SundusCase(dim=dim, k=k, n=2 if dim == 3 else 3)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_sundus_reassembly_reuses_the_plan():
    """Newton / time loop usage: zero the system, change the iterate, scatter again -- one plan, new values."""
    from nfempy_b200.assembly import GPUAssembler
    from tests.parity_helpers import SundusCase

    case = SundusCase(dim=3, k=1, n=2)
    space = case.space()
    a_n, a = case.alphas(space.n_dof)
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    asm.form_to_input_list = case.gpu_forms(space, (a_n, a))
    first = [v.copy() for v in asm.assemble()]
    asm.form_to_input_list = case.gpu_forms(space, (a_n + 0.1, a))
    asm.jacobian.scale(0.0)
    asm.residual[:] = 0.0
    asm.scatter_forms()
    assert len(asm._plans) == 1
    second = asm.jacobian.getValuesCSR()
    assert np.array_equal(first[1], second[1]) and not np.array_equal(first[2], second[2])
    asm.form_to_input_list = case.gpu_forms(space, (a_n, a))
    again = asm.assemble()
    assert np.array_equal(again[2], first[2]) and np.array_equal(again[3], first[3])


@pytest.mark.parametrize("kind", ["degenerate", "unscaled", "projection"])
@pytest.mark.parametrize("dim,k", [(3, 1), (3, 2), (2, 2)])
def test_degenerate_elliptic_family(kind, dim, k):
    """DegenerateEllipticWeakForm / UnscaledEllipticWeakForm (+ their natural BC forms) and
    ScaledToPhysicalL2Projection with a porosity that vanishes on part of the domain, nonzero iterate."""
    from tests.parity_helpers import degenerate_case

    case = degenerate_case(kind=kind, dim=dim, k=k, n=2 if dim == 3 else 3)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_scaled_to_physical_evaluate_form_signature():
    """The reference's three-argument ``evaluate_form(i, alpha, alpha_scaled)`` on one cell vs the oracle."""
    from oracle import ref_forms
    from tests.parity_helpers import degenerate_case

    case = degenerate_case(kind="projection", dim=3, k=1, n=2)
    space, rspace = case.space(), case.ref_space()
    form = case.gpu_forms(space, None)["volume"][2]
    nl = len(space.destination_indexes(3))
    a, s = np.cos(np.arange(nl) * 0.7), np.sin(np.arange(nl) * 0.3) + 0.2
    r, j = form.evaluate_form(3, a, s)
    r0, j0 = ref_forms.scaled_to_physical_l2_projection(rspace, form.functions, 3, a, s)
    assert np.abs(r - r0).max() <= 1e-12 * np.abs(r0).max() and np.abs(j - j0).max() <= 1e-12 * np.abs(j0).max()


@pytest.mark.parametrize("dim,k,family", [(3, 2, "RT"), (3, 3, "RT"), (2, 3, "RT"), (3, 2, "BDM"), (3, 3, "BDM")])
def test_mass_block_on_the_fp64_tensor_cores(dim, k, family, monkeypatch):
    """NFA_HDIV_MMA=1: the symmetric flux mass block as 8 x 8 DMMA tiles (mma.sync m8n8k4) -- same matrix."""
    case = mixed_poisson_case(dim=dim, k=k, n=1 if dim == 3 else 2, family=family, varying=True)
    ref = case.run_oracle()
    monkeypatch.setenv("NFA_HDIV_MMA", "1")
    compare_csr(case.run_gpu(), ref, RTOL)
