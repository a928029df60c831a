#!/usr/bin/env python
"""BASELINE config 5 (two-flux transport, RT1 x RT1 x P0 x P0, repeated reassembly on one plan) on N GPUs of one node:
contiguous cell ranges, the halo layer replicated instead of an exchange (``HaloReplicatedAssembler``), every rank's row
block final without a data-path collective; the row-owner kernels (csrc/nfa_hdivrows.cu) do the work.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config5_sharded.py --box 69 --mode strong
  --mode strong: one mesh of box^3 x 6 tets split over the ranks; --mode weak: box^3 x 6 tets PER RANK (the box is
  stretched along x, so the replicated numbering on the host grows with N).

Timing: W warm-ups, K reassemblies bracketed by a barrier + synchronize, CUDA events, MAX over ranks; rank 0 prints one
JSON line.  Check that does not depend on the size: on every rank the owned rows of 64 sampled own cells are compared
with a single-GPU style assembly of just those cells and their neighbours (bit-identical)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--box", type=int, default=69)
    ap.add_argument("--mode", choices=("strong", "weak"), default="strong")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.assembly.sharding import HaloReplicatedAssembler
    from nfempy_b200.mesh import box_tet_mesh
    from tests.parity_helpers import SundusCase

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    nx = args.box * world if args.mode == "weak" else args.box
    mesh = box_tet_mesh(args.box, nx=nx, jitter=0.2)
    case = SundusCase(dim=3, k=1, mesh=mesh)
    space = case.space()
    t_space = time.perf_counter() - t0
    alphas = SundusCase.alphas(space.n_dof)
    h = HaloReplicatedAssembler(space, world, rank).set_forms(case.gpu_forms(space, alphas))
    asm = h.asm
    t0 = time.perf_counter()
    state = asm.prepare()
    torch.cuda.synchronize()
    t_prepare = time.perf_counter() - t0
    plan = state[4]
    rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=state[0].device)
    # fresh="rows": the rank's kernels replace the right-hand-side rows its cells feed; the other rows of the global-sized
    # buffer are never swept (a rank's step must not grow with the number of ranks)
    graph = asm.capture(state, rhs, fresh="rows")

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        graph.replay()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        graph.replay()
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    n_cells_rank = torch.tensor([float(len(h.cells)), float(h.cell_range[1] - h.cell_range[0])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(n_cells_rank) for _ in range(world)]
        dist.all_gather(gathered, n_cells_rank)
    else:
        gathered = [n_cells_rank]
    # size-independent check: owned rows of sampled own cells against an assembly of just those cells + neighbours
    indptr = plan.indptr.cpu().numpy()
    dest = np.asarray(space.destination_table())
    c0, c1 = h.cell_range
    sample = np.unique(np.linspace(c0, c1 - 1, 64).astype(np.int64))
    owned_mask = np.zeros(space.n_dof, dtype=bool)
    owned_mask[h.owned_rows] = True
    rows = np.unique(dest[sample].ravel())
    rows = rows[owned_mask[rows]]
    feeds = np.flatnonzero(np.isin(dest, rows).any(axis=1))
    small = GPUAssembler(space, None, np.zeros(space.n_dof))
    forms = case.gpu_forms(space, alphas)
    forms["volume"][1] = feeds
    forms.pop("bc")
    small.form_to_input_list = forms
    s_indptr, s_indices, s_vals, s_rhs = small.assemble()
    # the sampled rows carry no boundary terms in `small`; compare the matrix rows (boundary forms add zeros there)
    vals = asm.jacobian.values_dev.cpu().numpy()
    indices = plan.indices.cpu().numpy()
    ok = True
    for r in rows:
        a, b, sa, sb = indptr[r], indptr[r + 1], s_indptr[r], s_indptr[r + 1]
        ok &= bool(np.array_equal(indices[a:b], s_indices[sa:sb]) and np.array_equal(vals[a:b], s_vals[sa:sb]))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        n_total = len(dest)
        assembled = [int(g[0].item()) for g in gathered]
        print(json.dumps({
            "case": "config5 sharded: Sundus RT1 x RT1 x P0 x P0, halo-replicated partition (no exchange), graph replay",
            "mode": args.mode, "n_gpus": world, "cells_total": n_total, "cells_assembled_per_rank": assembled,
            "halo_fraction": sum(assembled) / n_total - 1.0, "n_dof": int(space.n_dof), "path": "row owner" if plan.rows else "two-stage",
            "ms_per_reassembly_max_over_ranks": float(ms.item()), "elements_per_s": n_total / (float(ms.item()) * 1e-3),
            "steps": args.steps, "warmup": args.warmup, "sampled_rows_bit_identical": bool(flag.item()),
            "rows_checked_rank0": int(len(rows)), "host_mesh_space_s": t_space, "prepare_s": t_prepare, "dtype": "f64",
            "scaling": args.mode,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
