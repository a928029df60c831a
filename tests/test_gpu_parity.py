"""GPU parity tests (run on the B200 box with -m gpu): CUDA path through the C ABI vs the cell-by-cell oracle.

Bar: CSR pattern bit-exact (int32 indptr/indices identical), FP64 values within 1e-12 of the oracle relative to
max(|value|, largest |value| in its row); see tests/parity_helpers.compare_csr.
"""
import numpy as np
import pytest

from tests.parity_helpers import compare_csr, elasticity_case, laplace_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("with_bc", [False, True])
def test_elasticity_p2_3d(fused, with_bc):
    case = elasticity_case(n=2, with_bc=with_bc)
    compare_csr(case.run_gpu(allow_fused=fused), case.run_oracle(), RTOL)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("kind", ["neumann", "normal"])
def test_elasticity_mixed_boundary_forms_on_subsequences(fused, kind):
    """Dirichlet penalty on every other boundary facet, traction load or normal-only penalty on the rest: exercises
    explicit (non-range) sequences, three segments in one plan, and the remaining two boundary forms."""
    case = elasticity_case(n=2, with_bc=True, neumann=(kind == "neumann"), normal_bc=(kind == "normal"))
    compare_csr(case.run_gpu(allow_fused=fused), case.run_oracle(), RTOL)


def test_elasticity_p2_3d_varying_coefficients():
    case = elasticity_case(n=2, with_bc=True, varying=True)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


@pytest.mark.parametrize("fused", [True, False])
def test_elasticity_p2_3d_nonzero_alpha(fused):
    case = elasticity_case(n=2, with_bc=True, alpha=True)
    compare_csr(case.run_gpu(allow_fused=fused), case.run_oracle(), RTOL)


@pytest.mark.parametrize("fused", [True, False])
def test_elasticity_p1_3d(fused):
    case = elasticity_case(n=2, with_bc=True, k=1)
    compare_csr(case.run_gpu(allow_fused=fused), case.run_oracle(), RTOL)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("k", [1, 2])
def test_elasticity_2d(fused, k):
    case = elasticity_case(n=3, with_bc=True, k=k, dim=2)
    compare_csr(case.run_gpu(allow_fused=fused), case.run_oracle(), RTOL)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_laplace(fused, k, dim):
    case = laplace_case(n=4 if dim == 2 else 2, dim=dim, k=k, with_bc=True)
    compare_csr(case.run_gpu(allow_fused=fused), case.run_oracle(), RTOL)


def test_laplace_varying_kappa_and_alpha():
    case = laplace_case(n=4, dim=2, k=2, with_bc=True, varying=True, alpha=True)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_evaluate_form_single_cell_matches_oracle():
    from nfempy_b200.weak_forms import LEPrimalWeakForm
    from oracle import ref_forms

    case = elasticity_case(n=1, with_bc=False, varying=True)
    space, rspace = case.space(), case.ref_space()
    wf = LEPrimalWeakForm(space)
    wf.functions = case.gpu_forms(space, None)["volume"][2].functions
    rng = np.random.default_rng(0)
    for iel in range(len(space.discrete_spaces["u"].elements)):
        alpha_l = rng.standard_normal(30)
        r, j = wf.evaluate_form(iel, alpha_l)
        r0, j0 = ref_forms.le_primal(rspace, wf.functions, iel, alpha_l)
        s = np.abs(j0).max()
        assert np.abs(j - j0).max() <= RTOL * s
        assert np.abs(r - r0).max() <= RTOL * max(np.abs(r0).max(), s)


def test_bitwise_deterministic():
    case = elasticity_case(n=3, with_bc=True, alpha=True)
    a = case.run_gpu(allow_fused=True)
    b = case.run_gpu(allow_fused=True)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    c = case.run_gpu(allow_fused=False)
    d = case.run_gpu(allow_fused=False)
    assert np.array_equal(c[2], d[2]) and np.array_equal(c[3], d[3])


def test_fused_equals_two_stage_midsize():
    """8 x 8 x 8 x 6 = 3072 tets: too slow for the oracle, so the two CUDA paths check each other + invariants."""
    import scipy.sparse as sp

    case = elasticity_case(n=8, with_bc=False)
    f = case.run_gpu(allow_fused=True)
    g = case.run_gpu(allow_fused=False)
    compare_csr(f, g, RTOL)
    A = sp.csr_matrix((f[2], f[1], f[0]))
    n = A.shape[0] // 3
    scale = np.abs(f[2]).max()
    assert abs(A - A.T).max() <= 1e-13 * scale
    for c in range(3):  # rigid translations are in the null space of the stiffness matrix
        t = np.zeros(3 * n)
        t[c::3] = 1.0
        assert np.abs(A @ t).max() <= 1e-12 * scale


def test_empty_sequence_and_reassembly_accumulates():
    from nfempy_b200.assembly import GPUAssembler

    case = elasticity_case(n=2, with_bc=True)
    space = case.space()
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    asm.form_to_input_list = case.gpu_forms(space, np.zeros(space.n_dof))
    ip, ix, d1, r1 = asm.assemble()
    d1, r1 = d1.copy(), r1.copy()
    asm.scatter_forms()  # second scatter without zeroing: PETSc ADD_VALUES semantics
    _, _, d2 = asm.jacobian.getValuesCSR()
    # (v + b) + v + b differs from 2 (v + b) by round-off only
    assert np.allclose(d2, 2.0 * d1, rtol=1e-14, atol=0.0)
    assert np.allclose(asm.residual, 2.0 * r1, rtol=1e-13, atol=1e-13 * np.abs(r1).max())
    asm.jacobian.scale(0.0)
    asm.residual[:] = 0.0
    asm.scatter_forms()
    assert np.array_equal(asm.jacobian.getValuesCSR()[2], d1)


def test_unknown_form_raises():
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.weak_forms import WeakForm

    class HostOnly(WeakForm):
        def evaluate_form(self, element_index, alpha):
            return None

    case = elasticity_case(n=1, with_bc=False)
    space = case.space()
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    asm.form_to_input_list = {"x": ["form", range(6), HostOnly(space), np.zeros(space.n_dof)]}
    with pytest.raises(NotImplementedError):
        asm.scatter_forms()


def test_example_driver_scatter_policy_primal_elasticity():
    """cg_elasticity_example_1.py:92-145: filtered + upper + explicit diagonal for the volume form, filtered + upper for
    the Dirichlet penalty form."""
    case = elasticity_case(n=2, with_bc=True, k=2)
    got = case.run_gpu(pattern_policy="filtered_upper", explicit_diagonal=True)
    ref = case.run_oracle(upper=True, explicit_diagonal=True)
    compare_csr(got, ref, RTOL)


def test_filtered_pattern_plan_is_reused_while_the_masks_do_not_change():
    """A Newton loop on the example drivers' value-dependent pattern: the second assembly finds the same masks and reuses
    the plan (no second sort of the kept entries); another load changes nothing in the pattern either; a coefficient that
    zeroes entries does, and the plan is rebuilt."""
    from nfempy_b200.assembly import GPUAssembler

    case = elasticity_case(n=2, with_bc=True, k=2)
    space = case.space()
    asm = GPUAssembler(space, None, np.zeros(space.n_dof), pattern_policy="filtered_upper", explicit_diagonal=True)
    asm.form_to_input_list = case.gpu_forms(space, None)
    first = [np.array(v, copy=True) for v in asm.assemble()]
    plan = asm._filtered_plan
    assert plan.reused == 0
    again = asm.assemble()
    assert asm._filtered_plan is plan and plan.reused == 1
    assert all(np.array_equal(a, b) for a, b in zip(first, again))
    compare_csr(again, case.run_oracle(upper=True, explicit_diagonal=True), RTOL)
    # other coefficients: whatever happens to the masks, the system is the one a fresh assembler builds
    wf = asm.form_to_input_list["volume"][2]
    fns = dict(wf.functions)
    fns["lambda"] = lambda x, y, z: 0.0 * x
    wf.functions = fns
    third = [np.array(v, copy=True) for v in asm.assemble()]
    fresh = GPUAssembler(space, None, np.zeros(space.n_dof), pattern_policy="filtered_upper", explicit_diagonal=True)
    fresh.form_to_input_list = asm.form_to_input_list
    want = fresh.assemble()
    assert all(np.array_equal(a, b) for a, b in zip(third, want))
    # another sequence: other destination tables -> the plan is rebuilt
    before = asm._filtered_plan
    asm.form_to_input_list["volume"][1] = list(range(3, 40))
    asm.assemble()
    assert asm._filtered_plan is not before and asm._filtered_plan.reused == 0


def test_scatter_forms_reuses_cached_coefficients_and_tracks_the_state_vector():
    """Second and later ``scatter_forms()`` calls (a Newton loop written the reference's way,
    R/examples/geothermal_flow/three_fields_advection_diffusion.py:397-431) reuse the evaluated coefficient arrays and the
    plan, only refresh the state vector, and give the same system as a fresh assembler; rebinding ``functions`` is seen."""
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    case = elasticity_case(n=2, with_bc=True, alpha=True)
    space = case.space()
    n = space.n_dof
    rng = np.random.default_rng(3)
    calls = {"n": 0}

    def lam(x, y, z):
        calls["n"] += 1
        return 2.0 * np.ones_like(x)

    form = LEPrimalWeakForm(space)
    form.functions = {"rhs": lambda x, y, z: np.array([x, y * z, 1.0 + 0 * x]), "lambda": lam, "mu": lambda x, y, z: np.ones_like(x)}
    n_cells = len(space.discrete_spaces["u"].elements)

    def fresh(alpha, functions):
        f2 = LEPrimalWeakForm(space)
        f2.functions = functions
        a2 = GPUAssembler(space, None, np.zeros(n), cache_coefficients=False)
        a2.form_to_input_list = {"v": ["form", range(n_cells), f2, alpha]}
        return a2.assemble()

    asm = GPUAssembler(space, None, np.zeros(n))
    alpha = np.zeros(n)  # first iterate: zero state
    asm.form_to_input_list = {"v": ["form", range(n_cells), form, alpha]}
    for it in range(3):
        asm.jacobian.scale(0.0)
        asm.residual[:] = 0.0
        asm.scatter_forms()
        got = asm.jacobian.getValuesCSR() + (asm.residual.copy(),)
        ref = fresh(alpha.copy(), form.functions)
        compare_csr(got, ref, 1e-13)
        alpha += rng.standard_normal(n)  # mutated in place, like alpha += delta in the reference's loop
    assert calls["n"] == 1 + 3, "the cached assembler evaluates lambda once (the three fresh ones once each)"
    # a rebound coefficient is picked up
    form.functions = dict(form.functions, **{"lambda": lambda x, y, z: 5.0 * np.ones_like(x)})
    asm.jacobian.scale(0.0)
    asm.residual[:] = 0.0
    asm.scatter_forms()
    compare_csr(asm.jacobian.getValuesCSR() + (asm.residual.copy(),), fresh(alpha.copy(), form.functions), 1e-13)


def test_captured_graph_follows_a_state_vector_that_starts_at_zero():
    """a graph captured while alpha == 0 must still carry the K.alpha term once alpha is refreshed."""
    import torch

    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    case = elasticity_case(n=2, with_bc=False)
    space = case.space()
    n = space.n_dof
    form = LEPrimalWeakForm(space)
    form.functions = {"rhs": lambda x, y, z: np.array([x, y, z]), "lambda": lambda x, y, z: np.ones_like(x),
                      "mu": lambda x, y, z: np.ones_like(x)}
    n_cells = len(space.discrete_spaces["u"].elements)
    asm = GPUAssembler(space, None, np.zeros(n))
    alpha = np.zeros(n)
    asm.form_to_input_list = {"v": ["form", range(n_cells), form, alpha]}
    state = asm.prepare()
    rhs = torch.zeros(n, dtype=torch.float64, device=state[0].device)
    graph = asm.capture(state, rhs)
    graph.replay()
    torch.cuda.synchronize()
    r0 = rhs.cpu().numpy().copy()
    alpha[:] = np.random.default_rng(1).standard_normal(n)
    asm.prepare()  # refreshes the persistent device state vector in place
    graph.replay()
    torch.cuda.synchronize()
    r1 = rhs.cpu().numpy()
    K = asm.jacobian.to_scipy()
    assert np.abs(r1 - (r0 + K @ alpha)).max() <= 1e-12 * np.abs(r1).max()
    assert np.abs(r1 - r0).max() > 1e-3


@pytest.mark.parametrize("with_bc", [False, True])
def test_right_hand_side_paths_agree(monkeypatch, with_bc):
    """The fused form's right-hand side three ways: tile-merged load (nfa_bplan_load_rhs: 750 tets = 11 full tiles of 64
    cells + a ragged one), local load vectors + row-owner gather (nfa_bplan_gather_load), the plan's generic row lists."""
    case = elasticity_case(n=5, with_bc=with_bc)
    merged = case.run_gpu(allow_fused=True)
    monkeypatch.setenv("NFA_LOAD_MERGE", "0")
    gathered = case.run_gpu(allow_fused=True)
    monkeypatch.setenv("NFA_RHS_GATHER", "0")
    generic = case.run_gpu(allow_fused=True)
    assert np.array_equal(gathered[3], generic[3])  # both sum cell by cell in ascending order
    scale = np.abs(generic[3]).max()
    assert np.abs(merged[3] - generic[3]).max() <= 1e-14 * scale  # associated tile by tile
    assert np.array_equal(merged[2], generic[2])
