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


@pytest.mark.parametrize("n,cap,riffle", [(1, 448, 296), (2, 448, 296), (2, 64, 3), (3, 200, 2)])
def test_item_plan_reproduces_the_block_sums(n, cap, riffle):
    mesh, space, s, node_dest, n_nodes, cell_nodes = _setup(n)
    bptr, bcols, keys = _block_csr(node_dest, n_nodes)
    rng = np.random.default_rng(5)
    lam, mu = 1.0 + rng.random(len(node_dest)), 0.5 + rng.random(len(node_dest))
    plan = p2_items.build(torch.as_tensor(node_dest), torch.as_tensor(bptr), torch.as_tensor(bcols), n_nodes, cap=cap, riffle=riffle)
    assert plan["n_entries"] == 52 * len(node_dest)
    rec = p2_items.geometry_records(mesh.points, cell_nodes, lam, mu)
    got = p2_items.walk(plan, rec, len(keys))
    assert not np.isnan(got).any(), "some CSR block was never written"
    ref = _brute_force(space, s, mesh, node_dest, n_nodes, cell_nodes, keys, bptr, lam, mu)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    # the staged kernel's chunk-local lists (cell slots, relative offsets) describe the same sums, bit for bit
    assert np.array_equal(p2_items.walk(plan, rec, len(keys), staged=True), got)
    hdr = plan["chunk_hdr"].numpy()
    hdr = hdr[np.argsort(hdr[:, 0])]  # chunks tile the block range exactly once (they are listed in riffled order)
    assert hdr[0, 0] == 0 and hdr[:, 1].sum() == len(keys) and np.all(hdr[1:, 0] == hdr[:-1, 0] + hdr[:-1, 1])
    assert np.all(hdr[:, 1] <= cap + np.diff(bptr).max())
    # prefetch lists: exactly the cells the chunk's entries name
    ell = plan["ell"].numpy().astype(np.int64) & 0xFFFFFFFF
    cptr, ccells = plan["chunk_cells_ptr"].numpy(), plan["chunk_cells"].numpy()
    th = plan["task_hdr"].numpy().astype(np.int64)
    for ci, (_, _, task0, ntask) in enumerate(plan["chunk_hdr"].numpy()):
        seen = set()
        for t in range(task0, task0 + ntask):
            seg = ell[th[t, 1] : th[t, 1] + 32 * ((th[t, 0] >> 8) & 0xFF)]
            seen.update((seg[seg != 0xFFFFFFFF] >> 6).tolist())
        assert sorted(seen) == ccells[cptr[ci] : cptr[ci + 1]].tolist()


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
