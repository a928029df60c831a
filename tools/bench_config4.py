#!/usr/bin/env python
"""BASELINE config 4 at its stated order and (per GPU) size: LCEScaledDualWeakForm + natural boundary conditions on
BDM_3 x RT_3 x P_2^disc x P_2^disc (348 local dofs, 64 quadrature points) through the streamed assembler.

One JSON line: sizes, stage times (CUDA events around the local kernel and the scatter of every chunk), FP64 rate of the
local kernel against the measured DFMA peak, scatter bandwidth against the measured copy bandwidth, and two checks that
do not depend on the size: (a) sampled rows against local matrices recomputed for just their cells, (b) every row sum
against the scattered row sums of the local matrices.
Run on the GPU box:  python tools/bench_config4.py --box 23 > gpurun_out/config4.jsonl"""
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
    ap.add_argument("--box", dest="n", type=int, default=23, help="box of n^3 cubes, 6 tets each (23 -> 73 002 tets)")
    ap.add_argument("--chunk", type=int, default=2048)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--samples", type=int, default=12)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from nfempy_b200.assembly.streamed import StreamedAssembler, partition_cells
    from nfempy_b200.device import space_device
    from nfempy_b200.mesh import box_tet_mesh
    from tests.parity_helpers import tensor_mixed_case

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    t0 = time.perf_counter()
    mesh = box_tet_mesh(args.n, jitter=0.2, seed=7)
    case = tensor_mixed_case(kind="lce_scaled", dim=3, s_family="BDM", m_family="RT", k=args.k, mesh=mesh)
    space = case.space()
    t_space = time.perf_counter() - t0
    alpha = case.alpha_fn(space.n_dof)
    forms = case.gpu_forms(space, alpha)
    form = forms["volume"][2]
    dest = np.asarray(form.dest_table(), dtype=np.int64)
    cells, own = None, (0, len(dest))
    if world > 1:  # contiguous cell ranges, halo cells replicated instead of an exchange (assembly/streamed.py)
        cells, owned_rows, own = partition_cells(dest, world, rank)
    asm = StreamedAssembler(space, chunk_cells=args.chunk, cells=cells)
    asm.form_to_input_list = forms
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    csr, rhs = asm.assemble()
    wall = time.perf_counter() - t0
    st = dict(asm.stats)
    n_cells, n_loc = st["n_cells"], st["n_loc"]  # cells this rank assembled (own + halo)
    n_q = len(space.quadrature[3][1])
    dev = space_device(space, None)
    cell_ids = np.arange(len(dest)) if cells is None else cells

    # (a) sampled cells: all their rows against local matrices recomputed for the cell and its facet neighbours
    rng = np.random.default_rng(5)
    cells_of = {}
    sample = sorted(set(rng.integers(own[0], own[1], args.samples).tolist()) | {own[0], own[1] - 1})  # own cells: complete rows
    want_rows = np.unique(dest[sample].reshape(-1))
    near = cell_ids[np.isin(dest[cell_ids], want_rows).any(axis=1)] if cells is not None else None
    owners = near if near is not None else np.flatnonzero(np.isin(dest, want_rows).any(axis=1))  # cells touching those rows
    data = form.prepare(dev, owners)
    al = torch.as_tensor(alpha).to(dev.device)
    dsub = torch.as_tensor(dest[owners].astype(np.int32)).to(dev.device)
    Ke, _ = form.local_blocks(dev, data, al, dsub)
    Ke = Ke.cpu().numpy()
    worst = 0.0
    for r in want_rows[:: max(1, len(want_rows) // 400)]:
        ip, idx, vals = csr.rows(int(r), int(r) + 1)
        exp = {}
        for k, c in enumerate(owners):
            hit = np.flatnonzero(dest[c] == r)
            if len(hit):
                for b, col in enumerate(dest[c]):
                    exp[col] = exp.get(col, 0.0) + Ke[k, hit[0], b]
        cols = np.array(sorted(exp))
        assert np.array_equal(idx, cols), "row %d: column pattern differs" % r
        ref = np.array([exp[c] for c in cols])
        worst = max(worst, float(np.abs(vals - ref).max() / max(np.abs(ref).max(), 1e-300)))
    # (b) row sums of the whole matrix against the scattered row sums of the local matrices (second pass of local kernels)
    rowsum_ref = torch.zeros(csr.n_dof, dtype=torch.float64, device=dev.device)
    for c0 in range(0, n_cells, args.chunk):
        c1 = min(n_cells, c0 + args.chunk)
        d = form.prepare(dev, cell_ids[c0:c1])
        K, _ = form.local_blocks(dev, d, None, None)
        rowsum_ref.index_add_(0, torch.as_tensor(dest[cell_ids[c0:c1]].reshape(-1)).to(dev.device), K.sum(dim=2).reshape(-1))
        del K, d
    complete = torch.zeros(csr.n_dof, dtype=torch.bool, device=dev.device)  # rows the rank's OWN cells touch
    complete[torch.as_tensor(np.unique(dest[own[0] : own[1]].reshape(-1))).to(dev.device)] = True
    block = 1 << 20
    worst_sum = 0.0
    scale = float(rowsum_ref.abs().max().item())
    for r0 in range(0, csr.n_dof, block):
        r1 = min(csr.n_dof, r0 + block)
        a, b = int(csr.indptr[r0].item()), int(csr.indptr[r1].item())
        lengths = (csr.indptr[r0 + 1 : r1 + 1] - csr.indptr[r0:r1])
        s = torch.segment_reduce(csr.values[a:b], "sum", lengths=lengths)
        worst_sum = max(worst_sum, float(((s - rowsum_ref[r0:r1]).abs() * complete[r0:r1]).max().item()) / scale)
    dfma = 33.7  # measured DFMA peak of this pool's B200s (tools/ubench/fp64_peak.cu, profiles/r02_fp64_peak.jsonl)
    ppath = os.path.join(ROOT, "profiles", "r02_fp64_peak.jsonl")
    if os.path.exists(ppath):
        for ln in open(ppath):
            try:
                rec = json.loads(ln)
                if rec.get("kind", "").startswith("dfma"):
                    dfma = float(rec["tflops"])
            except Exception:
                pass
    hbm = 6548.2
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        hbm = float(json.load(open(mp))["hbm_gbs"])
    flops = 2.0 * n_q * n_loc * n_loc * n_cells
    tf = flops / (st["local_kernel_ms"] * 1e-3) / 1e12
    sc_bytes = 16.0 * n_loc * n_loc * n_cells
    own_cells = own[1] - own[0]
    if world > 1:
        t = torch.tensor([st["device_ms"], st["local_kernel_ms"], st["scatter_ms"], wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = torch.tensor([float(own_cells), float(n_cells), float(len(owned_rows)), worst, worst_sum], dtype=torch.float64, device="cuda")
        mx = tot.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        st["device_ms"], st["local_kernel_ms"], st["scatter_ms"], wall = (float(v) for v in (t[0], t[1], t[2], t[3] / 1e3))
        multi = {"n_gpus": world, "own_cells_total": int(tot[0].item()), "assembled_cells_total": int(tot[1].item()),
                 "halo_overhead": float(tot[1].item() / tot[0].item()) - 1.0, "owned_rows_total": int(tot[2].item()),
                 "partition": "contiguous cell ranges, halo cells replicated, no exchange; times = max over ranks",
                 "elements_per_s_device_all_gpus": float(tot[0].item()) / (st["device_ms"] * 1e-3)}
        worst, worst_sum = float(mx[3].item()), float(mx[4].item())
    else:
        multi = None
    if rank != 0:
        dist.destroy_process_group()
        return
    line = {
        "multi_gpu": multi,
        "case": "config 4: LCEScaledDualWeakForm + BC, BDM_%d x RT_%d x P_%d x P_%d, streamed" % (args.k, args.k, args.k - 1, args.k - 1),
        "cells": n_cells, "n_loc": n_loc, "n_q": n_q, "n_dof": st["n_dof"], "nnz": st["nnz"], "csr_gb": st["csr_bytes"] / 1e9,
        "chunk_cells": st["chunk_cells"], "n_chunks": st["n_chunks"], "host_space_s": t_space,
        "tables_and_pattern_s": st["tables_and_pattern_s"], "host_coefficients_s": st["host_coefficients_s"],
        "local_kernel_ms": st["local_kernel_ms"], "scatter_ms": st["scatter_ms"], "device_ms": st["device_ms"],
        "rhs_s": st["rhs_s"], "assemble_wall_s": wall,
        "elements_per_s_device": n_cells / (st["device_ms"] * 1e-3), "nonzeros_per_s_device": st["nnz"] / (st["device_ms"] * 1e-3),
        "roofline_local": {"bound": "fp64", "flops_per_cell": 2.0 * n_q * n_loc * n_loc, "achieved_tflops": tf,
                           "peak_tflops_dfma_measured": dfma, "frac": tf / dfma},
        "roofline_scatter": {"bound": "hbm", "bytes_per_cell": 16.0 * n_loc * n_loc,
                             "achieved_gbs": sc_bytes / (st["scatter_ms"] * 1e-3) / 1e9, "peak_gbs": hbm,
                             "frac": sc_bytes / (st["scatter_ms"] * 1e-3) / 1e9 / hbm},
        "check_sampled_rows_max_rel": worst, "check_sampled_cells": len(sample), "check_row_sums_max_rel": worst_sum,
        "table_bytes_per_cell": st["table_bytes"] / n_cells, "dtype": "f64", "rhs_norm": float(rhs.norm().item()),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
