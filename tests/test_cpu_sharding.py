"""CPU tests of the multi-GPU host logic: ownership / halo / interface lists of ``ShardPlan`` and the interface-row
exchange over ``torch.distributed`` (gloo, world_size 2 and 3) reproducing the single-process assembly."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nfempy_b200.assembly.sharding import ShardPlan, buffer_offsets, exchange_interface_rows
from nfempy_b200.mesh import box_tet_mesh
from nfempy_b200.spaces import ProductSpace


def _space(nx=6, n=2):
    mesh = box_tet_mesh(n, nx=nx, jitter=0.1)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {})
    return space


def _local_blocks(n_cells, n_loc, seed=7):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n_cells, n_loc, n_loc)), rng.standard_normal((n_cells, n_loc))


def _assemble(dest, Ke, re, n_dof, cells, weights=None):
    """CSR over the pattern of ``cells`` with the contributions of cells whose weight is 1 (halo cells weigh 0)."""
    d = dest[cells]
    w = np.ones(len(cells)) if weights is None else weights
    rows = np.repeat(d, d.shape[1], axis=1).ravel()
    cols = np.tile(d, (1, d.shape[1])).ravel()
    data = (Ke[cells] * w[:, None, None]).ravel()
    A = sp.coo_matrix((data, (rows, cols)), shape=(n_dof, n_dof)).tocsr()
    P = sp.coo_matrix((np.ones_like(data), (rows, cols)), shape=(n_dof, n_dof)).tocsr()
    A = A + 0 * P  # keep explicit zeros of the halo pattern
    A.sort_indices()
    full = (P != 0).astype(np.float64)
    full.sort_indices()
    vals = np.zeros(full.nnz)
    # align A's data onto the full pattern
    key_f = np.repeat(np.arange(n_dof), np.diff(full.indptr)).astype(np.int64) * n_dof + full.indices
    key_a = np.repeat(np.arange(n_dof), np.diff(A.indptr)).astype(np.int64) * n_dof + A.indices
    vals[np.searchsorted(key_f, key_a)] = A.data
    rhs = np.zeros(n_dof)
    np.add.at(rhs, d.ravel(), (re[cells] * w[:, None]).ravel())
    return full.indptr.astype(np.int32), full.indices.astype(np.int32), vals, rhs


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_shard_plan_invariants(world):
    space = _space()
    dest = space.destination_table()
    plans = [ShardPlan(dest, space.n_dof, world, r) for r in range(world)]
    owned = np.concatenate([p.owned_rows for p in plans])
    assert len(owned) == len(np.unique(owned)) == space.n_dof  # every dof has exactly one owner
    for r, p in enumerate(plans):
        assert p.is_own.sum() == len(p.own_cells)
        for q, rows in p.send_rows.items():
            assert q < r and np.array_equal(rows, plans[q].recv_rows[r])
        for q, rows in p.recv_rows.items():
            assert q > r and np.array_equal(rows, plans[q].send_rows[r])
        # every cell touching one of my rows is in my local cell list
        mask = np.zeros(space.n_dof, dtype=bool)
        mask[p.rows_touched] = True
        assert np.array_equal(p.local_cells, np.flatnonzero(mask[dest].any(axis=1)))
        # owned rows are a few contiguous intervals (vertex block + edge block of the numbering)
        assert 1 <= len(p.row_block_ranges()) <= 4


def _worker(rank, world, port, nx, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        space = _space(nx, n)
        dest = space.destination_table()
        n_dof = space.n_dof
        Ke, re = _local_blocks(len(dest), dest.shape[1])
        shard = ShardPlan(dest, n_dof, world, rank)
        indptr, indices, vals, rhs = _assemble(dest, Ke, re, n_dof, shard.local_cells, shard.is_own.astype(float))
        sizes = {}

        def make_buffer(tag, rows):
            off = buffer_offsets(indptr, rows)
            sizes[tag] = off
            return torch.zeros(int(off[-1]) + len(rows), dtype=torch.float64)

        def pack(rows, buf):
            off = buffer_offsets(indptr, rows)
            b = buf.numpy()
            for k, r in enumerate(rows):
                b[off[k] : off[k + 1]] = vals[indptr[r] : indptr[r + 1]]
            b[off[-1] :] = rhs[rows]

        def unpack_add(rows, buf):
            off = buffer_offsets(indptr, rows)
            b = buf.numpy()
            for k, r in enumerate(rows):
                vals[indptr[r] : indptr[r + 1]] += b[off[k] : off[k + 1]]
            rhs[rows] += b[off[-1] :]

        exchange_interface_rows(dist, shard, make_buffer, pack, unpack_add)
        # compare my owned rows with the single-process assembly
        g_indptr, g_indices, g_vals, g_rhs = _assemble(dest, Ke, re, n_dof, np.arange(len(dest)))
        ok = True
        for r in shard.owned_rows:
            a, b = indptr[r], indptr[r + 1]
            ga, gb = g_indptr[r], g_indptr[r + 1]
            ok &= np.array_equal(indices[a:b], g_indices[ga:gb])
            ok &= np.allclose(vals[a:b], g_vals[ga:gb], rtol=1e-13, atol=1e-13)
        ok &= np.allclose(rhs[shard.owned_rows], g_rhs[shard.owned_rows], rtol=1e-13, atol=1e-13)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        n_owned = torch.tensor([len(shard.owned_rows)])
        dist.all_reduce(n_owned)
        if rank == 0:
            out.put((int(flag.item()), int(n_owned.item()), n_dof))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_interface_exchange_over_gloo_matches_single_process(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6, 2, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ok, n_owned, n_dof = out.get(timeout=10)
    assert ok == 1 and n_owned == n_dof


# ---- halo-replicated partition (H(div) spaces: no exchange) -----------------------------------------------------------
def _sundus_space(n=3):
    from tests.parity_helpers import SundusCase

    case = SundusCase(dim=3, k=1, n=n)
    return case, case.space()


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_halo_replicated_partition_invariants(world):
    from nfempy_b200.assembly.sharding import HaloReplicatedAssembler

    case, space = _sundus_space()
    dest = space.destination_table()
    n_bc = len(space.discrete_spaces["mp"].bc_elements)
    owned_by, bc_by = np.full(space.n_dof, -1), np.full(n_bc, -1)
    for r in range(world):
        h = HaloReplicatedAssembler(space, world, r).set_forms(case.gpu_forms(space, case.alphas(space.n_dof)))
        assert (owned_by[h.owned_rows] == -1).all()
        owned_by[h.owned_rows] = r
        c0, c1 = h.cell_range
        assembled = np.zeros(len(dest), dtype=bool)
        assembled[h.cells] = True
        assert assembled[c0:c1].all()
        # every cell that feeds an owned row is assembled on the rank: the rows are complete without an exchange
        feeds = np.isin(dest, h.owned_rows).any(axis=1)
        assert assembled[feeds].all()
        kind, items, form = h.asm.form_to_input_list["bc"][:3]
        assert (bc_by[items] == -1).all()
        bc_by[items] = r
        assert np.array_equal(h.asm.form_to_input_list["volume"][1], h.cells)
    assert (owned_by >= 0).all() and (bc_by >= 0).all()


def _halo_worker(rank, world, port, out):
    from nfempy_b200.assembly.sharding import HaloReplicatedAssembler

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case, space = _sundus_space()
        dest = space.destination_table()
        n_dof = space.n_dof
        h = HaloReplicatedAssembler(space, world, rank)
        Ke, re = _local_blocks(len(dest), dest.shape[1])
        indptr, indices, vals, rhs = _assemble(dest, Ke, re, n_dof, h.cells)  # the rank's cells + halo, all weighing 1
        g_indptr, g_indices, g_vals, g_rhs = _assemble(dest, Ke, re, n_dof, np.arange(len(dest)))
        ok = True
        for r in h.owned_rows:
            a, b = indptr[r], indptr[r + 1]
            ga, gb = g_indptr[r], g_indptr[r + 1]
            ok &= np.array_equal(indices[a:b], g_indices[ga:gb]) and np.array_equal(vals[a:b], g_vals[ga:gb])
        ok &= np.array_equal(rhs[h.owned_rows], g_rhs[h.owned_rows])
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        owners = torch.zeros(n_dof, dtype=torch.int64)
        owners[torch.as_tensor(h.owned_rows)] = 1
        dist.all_reduce(owners)  # every row owned exactly once across the ranks
        if rank == 0:
            out.put((int(flag.item()), int(owners.min().item()), int(owners.max().item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_replicated_rows_are_final_without_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert out.get(timeout=10) == (1, 1, 1)


def test_halo_replicated_assembler_refuses_what_it_cannot_shard():
    from nfempy_b200.assembly.sharding import HaloReplicatedAssembler

    case, space = _sundus_space(2)
    forms = case.gpu_forms(space, case.alphas(space.n_dof))
    h = HaloReplicatedAssembler(space, 2, 0)
    part = dict(forms)
    part["volume"] = list(forms["volume"])
    part["volume"][1] = range(5)  # a volume form over part of the cells: the caller already partitioned, refuse
    with pytest.raises(NotImplementedError):
        h.set_forms(part)
    odd = dict(forms)
    odd["coupling"] = ["interface_form", [], forms["volume"][2], None]
    with pytest.raises(NotImplementedError):
        h.set_forms(odd)
    # world 1: everything is the rank's own
    h1 = HaloReplicatedAssembler(space, 1, 0).set_forms(forms)
    assert len(h1.cells) == len(space.destination_table()) and len(h1.owned_rows) == space.n_dof
    assert len(h1.asm.form_to_input_list["bc"][1]) == len(space.discrete_spaces["mp"].bc_elements)
