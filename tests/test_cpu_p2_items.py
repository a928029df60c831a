"""Row-item plan of the P2-tetrahedron fused kernel, walked on the CPU: every CSR block is produced exactly once and
equals the sum over cells of the reference-tensor formula the generic fused kernel uses."""
import numpy as np
import pytest
import torch

from nfempy_b200.assembly import p2_items
from nfempy_b200.mesh import box_tet_mesh
from nfempy_b200.spaces import ProductSpace
from nfempy_b200.weak_forms.h1_common import reference_stiffness_tensor


def _setup(n, jitter=0.2):
    mesh = box_tet_mesh(n, jitter=jitter, shuffle_local_nodes=True)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {"u": []})
    s = space.discrete_spaces["u"]
    dest = np.asarray(space.destination_table(), dtype=np.int64)
    node_dest = dest[:, ::3] // 3
    n_nodes = space.n_dof // 3
    cell_nodes = np.asarray(s.cell_nodes(), dtype=np.int64)
    return mesh, space, s, node_dest, n_nodes, cell_nodes


def _block_csr(node_dest, n_nodes):
    keys = np.unique((node_dest[:, :, None] * n_nodes + node_dest[:, None, :]).reshape(-1))
    rows, cols = keys // n_nodes, keys % n_nodes
    bptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(bptr, rows + 1, 1)
    return np.cumsum(bptr), cols, keys


def _brute_force(space, s, mesh, node_dest, n_nodes, cell_nodes, keys, bptr, lam, mu):
    points, weights = space.quadrature[3]
    M = reference_stiffness_tensor(s.ref_element, points, weights)
    X = mesh.points[cell_nodes]
    vals = np.zeros(len(keys) * 9)
    nb_row = np.diff(bptr)
    for c in range(len(X)):
        J = (X[c, 1:] - X[c, 0]).T
        det, Jinv = abs(np.linalg.det(J)), np.linalg.inv(J)
        for a in range(10):
            for b in range(10):
                T = Jinv.T @ M[a, b] @ Jinv
                K = det * (lam[c] * T + mu[c] * T.T + mu[c] * np.trace(T) * np.eye(3))
                N = node_dest[c, a]
                blk = np.searchsorted(keys, N * n_nodes + node_dest[c, b])
                base = 9 * bptr[N] + 3 * (blk - bptr[N])
                for r in range(3):
                    vals[base + r * 3 * nb_row[N] : base + r * 3 * nb_row[N] + 3] += K[r]
    return vals


@pytest.mark.parametrize("n,cap,riffle,optimise", [(1, 448, 296, False), (2, 448, 296, True), (2, 64, 3, True), (3, 200, 2, False)])
def test_item_plan_reproduces_the_block_sums(n, cap, riffle, optimise):
    mesh, space, s, node_dest, n_nodes, cell_nodes = _setup(n)
    bptr, bcols, keys = _block_csr(node_dest, n_nodes)
    rng = np.random.default_rng(5)
    lam, mu = 1.0 + rng.random(len(node_dest)), 0.5 + rng.random(len(node_dest))
    opt = (lambda gp, hdr: p2_items.greedy_lane_order(gp.numpy(), hdr.numpy(), window=8)) if optimise else None
    plan = p2_items.build(torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes, cap=cap,
                          riffle=riffle, lane_optimizer=opt)
    assert plan["n_entries"] == 46 * len(node_dest)
    rec = p2_items.geometry_records(mesh.points, cell_nodes, lam, mu)
    got = p2_items.walk(plan, rec, len(keys))
    assert not np.isnan(got).any(), "some CSR block was never written"
    # the items that own a node's diagonal block also own its right-hand-side rows (nfa_assemble_p2tet_elasticity_load)
    re4 = rng.standard_normal((len(node_dest), 10, 4))
    got2, rhs = p2_items.walk(plan, rec, len(keys), re4=re4, n_nodes=n_nodes)
    assert np.array_equal(got2, got) and not np.isnan(rhs).any(), "some right-hand-side row was never written"
    rhs_ref = np.zeros((n_nodes, 3))
    np.add.at(rhs_ref, node_dest.reshape(-1), re4[:, :, :3].reshape(-1, 3))
    assert np.abs(rhs.reshape(-1, 3) - rhs_ref).max() <= 1e-13 * np.abs(rhs_ref).max()
    assert plan["all_rows_nonempty"] == 1
    ref = _brute_force(space, s, mesh, node_dest, n_nodes, cell_nodes, keys, bptr, lam, mu)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    if optimise:  # the lane permutation moves items between lanes, never between sums: same values bit for bit
        plain = p2_items.build(torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes, cap=cap,
                               riffle=riffle)
        assert np.array_equal(p2_items.walk(plain, rec, len(keys)), got)
    hdr = plan["chunk_hdr"].numpy()
    hdr = hdr[np.argsort(hdr[:, 0])]  # chunks tile the block range exactly once (they are listed in riffled order)
    assert hdr[0, 0] == 0 and hdr[:, 1].sum() == len(keys) and np.all(hdr[1:, 0] == hdr[:-1, 0] + hdr[:-1, 1])
    assert np.all(hdr[:, 1] <= cap + np.diff(bptr).max())


def _store_wavefronts(plan):
    """Shared-memory wavefronts of the tile stores (STS.64: a half warp per wavefront, one per distinct bank pair) and
    their ideal count, from the item headers."""
    th = plan["task_hdr"].numpy().astype(np.int64)
    ih = plan["item_hdr"].numpy().astype(np.int64) & 0xFFFF
    total = ideal = 0
    for typ_k, _, slot0, nit in th:
        typ, nit = typ_k & 0xFF, nit & 63
        per = 2 if typ >= p2_items.TYPE_D else 1
        rows = ih[slot0 + per * np.arange(nit)].copy()
        if typ == p2_items.TYPE_A:  # the two unused block slots carry the row's node id
            rows[:, 1:3] = p2_items.NONE16
        extra = ih[slot0 + per * np.arange(nit) + 1][:, :2] if typ == p2_items.TYPE_E else np.full((nit, 2), p2_items.NONE16)
        sig = p2_items.store_signature(np.concatenate([rows, extra], axis=1))
        for h in range(0, nit, 16):
            for p in range(p2_items.N_SIG):
                b = sig[h : h + 16, p]
                b = b[b >= 0]
                if len(b):
                    total += 3 * np.bincount(b, minlength=16).max()
                    ideal += 3
    return total, ideal


def test_bank_aware_lane_order_removes_store_conflicts():
    _, _, _, node_dest, n_nodes, _ = _setup(4)
    bptr, bcols, _ = _block_csr(node_dest, n_nodes)
    args = (torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes)
    plain = p2_items.build(*args, cap=640)
    tuned = p2_items.build(*args, cap=640, lane_optimizer=lambda gp, hdr: p2_items.greedy_lane_order(gp.numpy(), hdr.numpy()))
    w0, i0 = _store_wavefronts(plain)
    w1, i1 = _store_wavefronts(tuned)
    assert w1 < 0.85 * w0, (w0, w1, i0, i1)


def test_item_plan_entries_are_in_ascending_cell_order():
    _, _, _, node_dest, n_nodes, _ = _setup(2)
    bptr, bcols, _ = _block_csr(node_dest, n_nodes)
    plan = p2_items.build(torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes)
    ell = plan["ell"].numpy().astype(np.int64) & 0xFFFFFFFF
    for typ_k, base, item0, nit in plan["task_hdr"].numpy():
        kmax, nit = (typ_k >> 8) & 0xFF, nit & 63
        for lane in range(nit):
            cells = [ell[base + 32 * k + lane] >> 6 for k in range(kmax) if ell[base + 32 * k + lane] != 0xFFFFFFFF]
            assert cells == sorted(cells) and len(set(cells)) == len(cells) and len(cells) >= 1


def test_balanced_split_covers_the_chunks_and_evens_out_the_work():
    _, _, _, node_dest, n_nodes, _ = _setup(3)
    bptr, bcols, _ = _block_csr(node_dest, n_nodes)
    plan = p2_items.build(torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes, cap=64, riffle=4)
    n_chunks = plan["n_chunks"]
    for n_cta in (1, 3, 7, n_chunks):
        split = p2_items.balanced_split(plan, n_cta).numpy()
        assert split[0] == 0 and split[-1] == n_chunks and len(split) == n_cta + 1 and np.all(np.diff(split) >= 0)
    th = plan["task_hdr"].numpy().astype(np.int64)
    ch = plan["chunk_hdr"].numpy().astype(np.int64)
    cost = np.array([(3 + ((th[t0 : t0 + nt, 0] >> 8) & 0xFF)).sum() for _, _, t0, nt in ch])
    split = p2_items.balanced_split(plan, 4).numpy()
    per = np.array([cost[split[b] : split[b + 1]].sum() for b in range(4)])
    assert per.max() - per.min() <= 2 * cost.max()  # within the granularity of whole chunks


def test_item_plan_without_the_entries_of_dead_cells():
    """``cell_live``: the halo cells of a rank's plan contribute zeros; their entries are cut from the lists, the items stay."""
    mesh, space, s, node_dest, n_nodes, cell_nodes = _setup(2)
    bptr, bcols, keys = _block_csr(node_dest, n_nodes)
    rng = np.random.default_rng(11)
    n = len(node_dest)
    lam, mu = 1.0 + rng.random(n), 0.5 + rng.random(n)
    live = rng.random(n) > 0.3
    live[:6] = False  # a whole cube of dead cells: some items keep no entry at all
    full = p2_items.build(torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes, cap=64, riffle=3)
    plan = p2_items.build(torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes, cap=64, riffle=3,
                          cell_live=torch.as_tensor(live))
    assert plan["n_items"] == full["n_items"] and plan["n_entries"] < full["n_entries"]
    ell = plan["ell"].numpy().astype(np.int64) & p2_items.NONE32
    used = ell[ell != p2_items.NONE32] >> 6
    assert live[used].all(), "an entry of a dead cell survived"
    rec = p2_items.geometry_records(mesh.points, cell_nodes, lam, mu)
    got = p2_items.walk(plan, rec, len(keys))
    assert not np.isnan(got).any(), "some CSR block was never written"
    ref = _brute_force(space, s, mesh, node_dest, n_nodes, cell_nodes, keys, bptr, np.where(live, lam, 0.0), np.where(live, mu, 0.0))
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
