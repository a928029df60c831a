"""GPU test of the cell-partitioned assembly: two (and three) ranks' local fused assemblies, interface rows packed /
added with the library kernels, owned row blocks compared with the single-GPU assembly.  Runs the ranks one after
the other on one device with a loop-back transfer (the NCCL leg itself is exercised by bench.py --gpus N and
tests/test_cpu_sharding.py covers the P2P protocol over gloo)."""
import numpy as np
import pytest

from nfempy_b200.mesh import box_tet_mesh
from nfempy_b200.spaces import ProductSpace

pytestmark = pytest.mark.gpu


def _f(x, y, z):
    return np.array([np.sin(x + y), z * x, np.cos(y) + 1.0])


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("varying", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_assembly_matches_single_gpu(world, varying, graph):
    """``varying``: cell-wise lambda (halo cells get zero coefficients) instead of mesh constants (halo cells switched off
    by the plan's cell mask, constant-rho kernel); ``graph``: local assembly replayed from a CUDA graph."""
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.assembly.sharding import ShardedAssembler
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    mesh = box_tet_mesh(3, nx=6, jitter=0.2)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {})
    n_cells = len(space.discrete_spaces["u"].elements)
    lam_cell = 2.0 + np.random.default_rng(7).random(n_cells)

    class CellwiseForm(LEPrimalWeakForm):  # lambda given per cell (no load): the general-rho instantiation
        def prepare(self, dev, items=None, arrays=None):
            sel = lam_cell if items is None else lam_cell[np.asarray(items)]
            return super().prepare(dev, items, arrays={"lambda": sel, "mu": np.array([1.0]), "rhs": None})

    form = (CellwiseForm if varying else LEPrimalWeakForm)(space)
    form.functions = {"rhs": _f, "lambda": lambda x, y, z: 2.0 * np.ones_like(x), "mu": lambda x, y, z: np.ones_like(x)}
    ref = GPUAssembler(space, None, np.zeros(space.n_dof))
    ref.form_to_input_list = {"volume": ["form", range(n_cells), form, None]}
    g_indptr, g_indices, g_vals, g_rhs = ref.assemble()

    ranks = [ShardedAssembler(space, form, world, r).prepare() for r in range(world)]
    for sa in ranks:
        if graph:
            sa.capture_local()
        sa.local_assemble()
    sent = {(r, q): sa.pack(q).clone() for r, sa in enumerate(ranks) for q in sa.shard.send_rows}
    for q, sa in enumerate(ranks):
        for r in sorted(sa.shard.recv_rows):
            sa.unpack_add(r, sent[(r, q)])
    owned_total = 0
    for sa in ranks:
        indptr = sa.plan.indptr.cpu().numpy()
        indices = sa.plan.indices.cpu().numpy()
        vals, rhs = sa.vals.cpu().numpy(), sa.rhs.cpu().numpy()
        rows = sa.shard.owned_rows
        owned_total += len(rows)
        scale = np.abs(g_vals).max()
        for r in rows:
            a, b, ga, gb = indptr[r], indptr[r + 1], g_indptr[r], g_indptr[r + 1]
            assert np.array_equal(indices[a:b], g_indices[ga:gb])
            assert np.abs(vals[a:b] - g_vals[ga:gb]).max() <= 1e-13 * scale
        assert np.abs(rhs[rows] - g_rhs[rows]).max() <= 1e-13 * max(np.abs(g_rhs).max(), 1.0)
    assert owned_total == space.n_dof


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_peer_memory_exchange_loopback(world):
    """nfa_push_rows / nfa_pull_add_rows with the ranks' receive arenas wired directly (one process, one device, no CUDA
    IPC): two consecutive steps exercise the flag / acknowledgement epochs; owned rows equal the single-GPU assembly."""
    from nfempy_b200 import _lib
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.assembly.sharding import ShardedAssembler
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    mesh = box_tet_mesh(3, nx=6, jitter=0.2)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {})
    n_cells = len(space.discrete_spaces["u"].elements)
    form = LEPrimalWeakForm(space)
    form.functions = {"rhs": _f, "lambda": lambda x, y, z: 2.0 * np.ones_like(x), "mu": lambda x, y, z: np.ones_like(x)}
    ref = GPUAssembler(space, None, np.zeros(space.n_dof))
    ref.form_to_input_list = {"volume": ["form", range(n_cells), form, None]}
    g_indptr, g_indices, g_vals, g_rhs = ref.assemble()
    ranks = [ShardedAssembler(space, form, world, r).prepare() for r in range(world)]
    arenas = [sa._p2p_alloc() for sa in ranks]
    for r, sa in enumerate(ranks):
        sa._p2p_wire(arenas[r], {q: arenas[q]["base"] for q in range(world)}, {q: arenas[q]["table"] for q in range(world)})
    for _ in range(2):
        for sa in reversed(ranks):  # one stream: a rank's pulls need the higher ranks' pushes to be queued before them
            sa.assemble()
    import torch

    torch.cuda.synchronize()
    assert _lib.load().nfa_p2p_timeouts() == 0
    scale = np.abs(g_vals).max()
    for sa in ranks:
        indptr = sa.plan.indptr.cpu().numpy()
        vals, rhs = sa.vals.cpu().numpy(), sa.rhs.cpu().numpy()
        for r in sa.shard.owned_rows:
            a, b, ga, gb = indptr[r], indptr[r + 1], g_indptr[r], g_indptr[r + 1]
            assert np.abs(vals[a:b] - g_vals[ga:gb]).max() <= 1e-13 * scale
        rows = sa.shard.owned_rows
        assert np.abs(rhs[rows] - g_rhs[rows]).max() <= 1e-13 * max(np.abs(g_rhs).max(), 1.0)
    for a in arenas:
        import ctypes

        _lib.load().nfa_p2p_free(ctypes.c_void_p(a["base"]))


@pytest.mark.parametrize("which,world", [("sundus", 3), ("poisson", 2)])
def test_halo_replicated_ranks_reproduce_the_single_gpu_rows(which, world):
    """H(div) spaces shard without an exchange (``HaloReplicatedAssembler``): every rank's OWNED rows -- pattern, values,
    right-hand side -- are bit-identical to the single-GPU assembly of the whole mesh, and the ranks' rows tile the
    system.  The ranks run one after the other on this device; the row-owner kernels do the work."""
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.assembly.sharding import HaloReplicatedAssembler
    from nfempy_b200.mesh import box_tet_mesh
    from tests.parity_helpers import SundusCase, mixed_poisson_case

    mesh = box_tet_mesh(4, jitter=0.2, shuffle_local_nodes=True)
    case = SundusCase(dim=3, k=1, mesh=mesh) if which == "sundus" else mixed_poisson_case(dim=3, k=1, mesh=mesh, varying=True)
    space = case.space()
    alphas = SundusCase.alphas(space.n_dof) if which == "sundus" else None
    whole = GPUAssembler(space, None, np.zeros(space.n_dof))
    whole.form_to_input_list = case.gpu_forms(space, alphas)
    g_indptr, g_indices, g_vals, g_rhs = [np.array(v, copy=True) for v in whole.assemble()]
    seen = np.zeros(space.n_dof, dtype=int)
    for rank in range(world):
        h = HaloReplicatedAssembler(space, world, rank).set_forms(case.gpu_forms(space, alphas))
        indptr, indices, vals, rhs, owned = h.assemble()
        assert list(h.asm._plans.values())[0].rows is not None  # the row-owner path, on part of the cells
        seen[owned] += 1
        for r in owned:
            a, b, ga, gb = indptr[r], indptr[r + 1], g_indptr[r], g_indptr[r + 1]
            assert np.array_equal(indices[a:b], g_indices[ga:gb]) and np.array_equal(vals[a:b], g_vals[ga:gb])
        assert np.array_equal(rhs[owned], g_rhs[owned])
    assert (seen == 1).all()
