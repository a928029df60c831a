"""GPU parity of the row-owner path for lowest-order H(div) x L2 forms (``csrc/nfa_hdivrows.cu``): mixed Poisson on
RT_1 x P_0 (config 3, k = 1) and the Sundus two-flux form on RT_1 x RT_1 x P_0 x P_0 (config 5) against the ORACLE, against
the two-stage path, on a non-zero matrix, from a replayed CUDA graph, and with more rows than one CTA pass."""
import os

import numpy as np
import pytest

from tests.parity_helpers import SundusCase, compare_csr, mixed_poisson_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _assembler(case, forms_alpha, **kw):
    from nfempy_b200.assembly import GPUAssembler

    space = case.space()
    asm = GPUAssembler(space, None, np.zeros(space.n_dof), **kw)
    asm.form_to_input_list = case.gpu_forms(space, forms_alpha(space.n_dof) if callable(forms_alpha) else forms_alpha)
    return space, asm


def _took_rows(asm):
    plans = list(asm._plans.values())
    return len(plans) == 1 and plans[0].rows is not None


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("with_bc,varying", [(True, False), (False, True), (True, True)])
def test_mixed_poisson_rt1_by_row_owner(dim, with_bc, varying):
    case = mixed_poisson_case(dim=dim, k=1, n=3, with_bc=with_bc, varying=varying)
    space, asm = _assembler(case, None)
    got = asm.assemble()
    assert _took_rows(asm)
    compare_csr(got, case.run_oracle(), RTOL)


@pytest.mark.parametrize("dim", [2, 3])
def test_sundus_rt1_by_row_owner(dim):
    case = SundusCase(dim=dim, k=1, n=3)
    space, asm = _assembler(case, SundusCase.alphas)
    got = asm.assemble()
    assert _took_rows(asm)
    compare_csr(got, case.run_oracle(), RTOL)


def test_sundus_with_zero_and_missing_state():
    """alpha = 0 for both levels (the first Newton iteration of a run that starts from rest)."""
    case = SundusCase(dim=3, k=1, n=2)
    zero = lambda n: (np.zeros(n), np.zeros(n))
    space, asm = _assembler(case, zero)
    got = asm.assemble()
    assert _took_rows(asm)
    space_r = case.ref_space()
    from oracle import ref_assembly, ref_forms

    mp = space_r.discrete_spaces["mp"]
    forms = [
        ("time_dependent_form", ref_forms.sundus_dual, case.functions, len(mp.element_ids)),
        ("time_dependent_bc_form", ref_forms.sundus_dual_bc_dirichlet, case.bc_functions, len(mp.bc_element_ids)),
    ]
    want = ref_assembly.assemble_time_dependent(space_r, forms, zero(space_r.n_dof), case.t)
    compare_csr(got, want, RTOL)


@pytest.mark.parametrize("which", ["sundus", "poisson"])
def test_row_owner_equals_two_stage(which, monkeypatch):
    """The same list through both paths: identical pattern, values and right-hand side within round-off -- on a mesh with
    more rows than one CTA pass of either kernel (6 x 6 x 6 x 6 tetrahedra: 1 296 cells, > 2 800 facets)."""
    from nfempy_b200.mesh import box_tet_mesh

    mesh = box_tet_mesh(6, jitter=0.2, shuffle_local_nodes=True)
    case = SundusCase(dim=3, k=1, mesh=mesh) if which == "sundus" else mixed_poisson_case(dim=3, k=1, mesh=mesh, varying=True)
    alphas = SundusCase.alphas if which == "sundus" else None
    _, asm = _assembler(case, alphas)
    rows = [np.array(v, copy=True) for v in asm.assemble()]
    assert _took_rows(asm)
    monkeypatch.setenv("NFA_HDIV_ROWS", "0")
    _, asm2 = _assembler(case, alphas)
    two = asm2.assemble()
    assert not _took_rows(asm2)
    compare_csr(rows, two, 1e-13)


def test_row_owner_adds_onto_existing_values_and_residual():
    """ADD_VALUES semantics across calls (R/src/assembly/ScatterFormData.py:5-25): a second scatter_forms() without
    scale(0.0) doubles matrix and residual."""
    case = SundusCase(dim=3, k=1, n=2)
    _, asm = _assembler(case, SundusCase.alphas)
    once = [np.array(v, copy=True) for v in asm.assemble()]
    asm.scatter_forms()
    twice = asm.jacobian.getValuesCSR()
    assert np.allclose(twice[2], 2.0 * once[2], rtol=1e-14, atol=0.0)
    assert np.allclose(asm.residual, 2.0 * once[3], rtol=1e-14, atol=1e-300)


def test_row_owner_graph_replay_follows_the_state():
    """One captured graph, the iterate refreshed in the persistent device buffer between replays."""
    import torch

    case = SundusCase(dim=3, k=1, n=2)
    space, asm = _assembler(case, SundusCase.alphas)
    a_n, a = SundusCase.alphas(space.n_dof)
    state = asm.prepare()
    rhs = torch.zeros(space.n_dof, dtype=torch.float64, device=state[0].device)
    graph = asm.capture(state, rhs)
    assert _took_rows(asm)
    graph.replay()
    torch.cuda.synchronize()
    first = (asm.jacobian.values_dev.cpu().numpy().copy(), rhs.cpu().numpy().copy())
    want = case.run_oracle()
    assert np.allclose(first[0], want[2], rtol=1e-11, atol=1e-12 * np.abs(want[2]).max())
    assert np.allclose(first[1], want[3], rtol=1e-11, atol=1e-12 * np.abs(want[3]).max())
    # new iterate: the volume form's first state buffer
    data = state[2][0]
    data.alpha_dev[0].copy_(torch.as_tensor(a_n + 0.25))
    graph.replay()
    torch.cuda.synchronize()
    second = asm.jacobian.values_dev.cpu().numpy()
    assert not np.array_equal(first[0], second)
    _, asm2 = _assembler(case, lambda n: (a_n + 0.25, a))
    ref = asm2.assemble()
    assert np.array_equal(second, ref[2]) and np.array_equal(rhs.cpu().numpy(), ref[3])


def test_row_owner_is_deterministic():
    case = SundusCase(dim=3, k=1, n=3)
    _, asm = _assembler(case, SundusCase.alphas)
    one = [np.array(v, copy=True) for v in asm.assemble()]
    two = asm.assemble()
    assert np.array_equal(one[2], two[2]) and np.array_equal(one[3], two[3])


def test_row_owner_on_part_of_the_cells(monkeypatch):
    """A sequence over part of the cells (one rank's share of a partitioned mesh): rows outside the share stay empty or
    are fed by the boundary form only; same CSR as the two-stage path, also when added onto existing values."""
    from nfempy_b200.assembly import GPUAssembler

    case = SundusCase(dim=3, k=1, n=3)
    space = case.space()
    alphas = SundusCase.alphas(space.n_dof)

    def run():
        asm = GPUAssembler(space, None, np.zeros(space.n_dof))
        forms = case.gpu_forms(space, alphas)
        n_cells = len(space.discrete_spaces["mp"].elements)
        forms["volume"][1] = list(range(7, n_cells - 5))
        asm.form_to_input_list = forms
        once = [np.array(v, copy=True) for v in asm.assemble()]
        asm.scatter_forms()
        return asm, once, asm.jacobian.getValuesCSR()[2].copy(), asm.residual.copy()

    asm, once, twice_vals, twice_rhs = run()
    assert _took_rows(asm) and not list(asm._plans.values())[0].rows_cover_all
    monkeypatch.setenv("NFA_HDIV_ROWS", "0")
    asm2, ref, ref_twice_vals, ref_twice_rhs = run()
    assert not _took_rows(asm2)
    compare_csr(once, ref, 1e-13)
    assert np.allclose(twice_vals, ref_twice_vals, rtol=1e-13, atol=1e-13 * np.abs(ref_twice_vals).max())
    assert np.allclose(twice_rhs, ref_twice_rhs, rtol=1e-13, atol=1e-13 * np.abs(ref_twice_rhs).max())


def test_chunked_state_upload_and_residual_add(monkeypatch):
    """The host side of ``scatter_forms()`` for long vectors (threaded staging of the state uploads into pinned memory,
    chunked ``residual += rhs``), forced onto a small problem: same system as the oracle's."""
    from nfempy_b200.assembly import gpu_assembler

    monkeypatch.setattr(gpu_assembler, "_CHUNK", 37)
    monkeypatch.setenv("NFA_HOST_THREADS", "4")  # the pool is sized from the core count; force it on small boxes too
    case = SundusCase(dim=3, k=1, n=2)
    _, asm = _assembler(case, SundusCase.alphas)
    got = [np.array(v, copy=True) for v in asm.assemble()]
    assert len(asm._state_stage) == 2 and asm._state_stage[0] is not None
    compare_csr(got, case.run_oracle(), RTOL)
    again = asm.assemble()  # second call: staging buffers and events reused
    assert np.array_equal(again[2], got[2]) and np.array_equal(again[3], got[3])


def test_new_time_reuses_geometry_and_matches_a_fresh_assembler():
    """A time loop written the reference's way: same forms, same functions, new ``t`` in the list -- the volume form reuses
    its mapped points, the boundary form its facet set-up; only the time-dependent callables are evaluated again."""
    case = SundusCase(dim=3, k=1, n=2)
    space, asm = _assembler(case, SundusCase.alphas)
    asm.assemble()
    forms = asm.form_to_input_list
    for entry in forms.values():
        entry[4] = 0.7
    got = [np.array(v, copy=True) for v in asm.assemble()]
    case2 = SundusCase(dim=3, k=1, n=2, t=0.7)
    _, fresh = _assembler(case2, SundusCase.alphas)
    want = fresh.assemble()
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    compare_csr(got, case2.run_oracle(), RTOL)


@pytest.mark.parametrize("kind", ["degenerate", "unscaled"])
@pytest.mark.parametrize("dim,alpha", [(3, True), (2, True), (3, False)])
def test_general_elliptic_rt1_by_row_owner(kind, dim, alpha):
    """Config 3's other form classes at k = 1 (DegenerateEllipticWeakForm / UnscaledEllipticWeakForm + their natural
    boundary forms; porosity vanishing on part of the domain) through the records kernel's form 2 and the rows kernel."""
    from tests.parity_helpers import degenerate_case

    case = degenerate_case(kind=kind, dim=dim, k=1, n=3, alpha=alpha)
    space, asm = _assembler(case, case.alpha_fn)
    got = asm.assemble()
    assert _took_rows(asm)
    compare_csr(got, case.run_oracle(), RTOL)


def test_projection_form_with_a_second_state_vector_keeps_the_two_stage_path():
    from tests.parity_helpers import degenerate_case

    case = degenerate_case(kind="projection", dim=3, k=1, n=2)
    space, asm = _assembler(case, case.alpha_fn)
    got = asm.assemble()
    assert not _took_rows(asm)
    compare_csr(got, case.run_oracle(), RTOL)
