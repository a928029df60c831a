"""Row-item kernel of vector-P2 elasticity on tetrahedra (nfa_assemble_p2tet_elasticity) against the generic fused
kernel (reference-tensor formulation) and the oracle: cell-wise varying lambda / mu, several chunk sizes, determinism."""
import os

import numpy as np
import pytest

from tests.parity_helpers import compare_csr, elasticity_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _assemble(n, lam, mu, p2tet, cap=None, seed_mesh=True, threads=None, grid=None, nbuf=None, laneopt=True):
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.mesh import box_tet_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    class ArrayForm(LEPrimalWeakForm):
        def prepare(self, dev, items=None, arrays=None):
            return super().prepare(dev, items, arrays={"lambda": lam, "mu": mu, "rhs": None})

    old = {k: os.environ.get(k) for k in ("NFA_P2TET", "NFA_P2TET_CAP", "NFA_P2TET_LANEOPT", "NFA_P2TET_THREADS", "NFA_P2TET_GRID", "NFA_P2TET_NBUF")}
    os.environ["NFA_P2TET"] = "1" if p2tet else "0"
    os.environ["NFA_P2TET_LANEOPT"] = "1" if laneopt else "0"
    os.environ.pop("NFA_P2TET_THREADS", None)
    if threads:
        os.environ["NFA_P2TET_THREADS"] = str(threads)
    for k, v in (("NFA_P2TET_GRID", grid), ("NFA_P2TET_NBUF", nbuf)):
        os.environ.pop(k, None)
        if v:
            os.environ[k] = str(v)
    if cap is not None:
        os.environ["NFA_P2TET_CAP"] = str(cap)
    try:
        mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True)
        space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
        space.build_structures({"u": [1]}, {"u": []})
        form = ArrayForm(space)
        form.functions = {}
        n_cells = len(space.discrete_spaces["u"].elements)
        asm = GPUAssembler(space, None, np.zeros(space.n_dof))
        asm.form_to_input_list = {"volume": ["form", range(n_cells), form, np.zeros(space.n_dof)]}
        out = asm.assemble()
        used = bool(getattr(next(iter(asm._plans.values())), "p2_stats", None))
        return [np.array(v) for v in out], used
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("laneopt,threads", [(True, None), (False, None), (True, 448)])
@pytest.mark.parametrize("n,cap", [(2, None), (5, None), (5, 64), (6, 1024)])
def test_row_item_kernel_matches_the_reference_tensor_kernel(n, cap, laneopt, threads):
    """``laneopt``: with / without the bank-aware lane order (nfa_p2_lane_order) -- items change lanes, never sums."""
    n_cells = 6 * n**3
    rng = np.random.default_rng(11)
    lam, mu = 0.5 + 3.0 * rng.random(n_cells), 0.2 + rng.random(n_cells)
    new, used_new = _assemble(n, lam, mu, True, cap, laneopt=laneopt, threads=threads)
    ref, used_ref = _assemble(n, lam, mu, False)
    assert used_new and not used_ref
    compare_csr(new, ref, RTOL)
    again, _ = _assemble(n, lam, mu, True, cap, laneopt=laneopt, threads=threads)
    assert np.array_equal(new[2], again[2]), "row-item kernel is not bit-reproducible"
    if not laneopt:
        other, _ = _assemble(n, lam, mu, True, cap, laneopt=True)
        assert np.array_equal(new[2], other[2]), "the lane order changed a sum"


@pytest.mark.parametrize("grid,nbuf,cap", [(1, 2, 64), (3, 2, 64), (2, 3, 100), (5, None, 48)])
def test_tile_hand_over_with_few_ctas(grid, nbuf, cap):
    """Few CTAs and few tiles: every CTA walks tens of chunks, so tiles are handed over many times (the bench-sized
    launch does this on every SM; the small meshes of the other tests give each CTA a single chunk)."""
    n = 6
    n_cells = 6 * n**3
    rng = np.random.default_rng(3)
    lam, mu = 0.5 + 3.0 * rng.random(n_cells), 0.2 + rng.random(n_cells)
    new, used = _assemble(n, lam, mu, True, cap, grid=grid, nbuf=nbuf)
    ref, _ = _assemble(n, lam, mu, False)
    assert used
    compare_csr(new, ref, RTOL)
    for _ in range(3):  # the relaxed tile-retire protocol must not lose a store: identical bits on every run
        again, _ = _assemble(n, lam, mu, True, cap, grid=grid, nbuf=nbuf)
        assert np.array_equal(new[2], again[2])


def test_row_item_kernel_constant_coefficients():
    """lambda, mu constant over the mesh (stride 0): the constant-rho instantiation."""
    n = 4
    new, used = _assemble(n, np.array([2.5]), np.array([0.75]), True, None)
    ref, _ = _assemble(n, np.array([2.5]), np.array([0.75]), False)
    assert used
    compare_csr(new, ref, RTOL)


def test_device_lane_order_equals_the_host_statement():
    """nfa_p2_lane_order (one device thread per group) against p2_items.greedy_lane_order on the same groups."""
    import torch

    from nfempy_b200 import _lib
    from nfempy_b200.assembly import p2_items

    rng = np.random.default_rng(0)
    sizes = [1, 5, 16, 17, 40, 130, 3]
    gp = np.concatenate([[0], np.cumsum(sizes)])
    n = int(gp[-1])
    hdr = np.full((n, 6), p2_items.NONE16, dtype=np.int64)
    hdr[:, 0] = 3 * rng.integers(0, 600, n)
    hdr[:, 3] = 3 * rng.integers(20, 70, n)
    two = rng.random(n) < 0.5
    hdr[two, 1] = 3 * rng.integers(0, 600, two.sum())
    hdr[two, 2] = 3 * rng.integers(0, 600, two.sum())
    hdr[two, 4] = 3 * rng.integers(0, 600, two.sum())
    ref = p2_items.greedy_lane_order(gp, hdr, window=24)
    lib = _lib.load()
    d_gp = torch.as_tensor(gp, dtype=torch.int64).cuda()
    d_h = torch.as_tensor(hdr, dtype=torch.int32).cuda().contiguous()
    perm = torch.empty(n, dtype=torch.int64, device="cuda")
    _lib.check(lib.nfa_p2_lane_order(len(sizes), _lib.ptr(d_gp), _lib.ptr(d_h), 24, _lib.ptr(perm), _lib.stream_ptr()), "lane order")
    assert np.array_equal(perm.cpu().numpy(), ref)


def test_row_item_kernel_against_the_oracle_with_boundary_forms():
    case = elasticity_case(n=2, with_bc=True, alpha=True)
    compare_csr(case.run_gpu(), case.run_oracle(), RTOL)


def test_non_positive_mu_falls_back_to_the_generic_kernel():
    n = 2
    lam, mu = np.ones(6 * n**3), np.ones(6 * n**3)
    mu[3] = 0.0
    new, _ = _assemble(n, lam, mu, True)
    ref, _ = _assemble(n, lam, mu, False)
    compare_csr(new, ref, RTOL)


@pytest.mark.parametrize("n_take", [47, 129])
def test_odd_cell_subsets_through_the_streaming_load_kernel(n_take):
    """An odd number of cells makes the last tile of the load stream end on an 8-byte (not 16-byte) boundary; the
    sub-sequence also exercises row items on a partially assembled mesh.  Checked against the oracle."""
    from nfempy_b200.assembly import GPUAssembler
    from oracle import ref_assembly, ref_forms

    case = elasticity_case(n=3, with_bc=False)
    space, rspace = case.space(), case.ref_space()
    forms = case.gpu_forms(space, np.zeros(space.n_dof))
    forms["volume"][1] = list(range(n_take))
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    asm.form_to_input_list = forms
    got = asm.assemble()
    kind, fn, functions, _ = case.ref_forms_fn(rspace)[0]
    ref = ref_assembly.assemble(rspace, [(kind, fn, functions, list(range(n_take)))], np.zeros(rspace.n_dof))
    # the GPU matrix keeps the full n_dof rows; rows never touched are empty in both
    compare_csr(got, ref, RTOL)
