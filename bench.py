#!/usr/bin/env python
"""Headline benchmark: elements/s assembled, 3-D linear elasticity, vector H1 p=2 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference ...                    # the reference algorithm on the host cores (oracle port)

Workload: unit cube, 55^3 cubes x 6 Kuhn tets = 998 250 tets per GPU, interior vertices jittered
(rng 1234), P2 x 3 displacement, lambda = mu = 1, smooth trigonometric load tabulated at the 27 quadrature points of
every cell (the API-faithful general load), alpha = 0.  One step = one complete assembly of the CSR values and the
right-hand side: geometry records -> fused local-contraction + deterministic global reduce -> load vector ->
right-hand-side reduce.  The sparsity plan is built once, outside the timed region, and reported separately.

At N > 1 the mesh is an N-times longer box (N x 998 250 tets); every rank owns a contiguous x-slab of cells, assembles
it with the same fused kernel (plus one halo layer of zero-weight cells) and the interface rows are summed onto their
owners by an NCCL send/recv inside the timed step (weak scaling; nfempy_b200/assembly/sharding.py).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--n", type=int, default=55, help="cubes per box edge (55 -> 998250 tets)")
    ap.add_argument("--cpu-cells", type=int, default=720, help="cells in the bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-varying", action="store_true", help="skip the cell-wise varying lambda, mu leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the captured CUDA graph")
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: exchange the interface rows with NCCL send/recv instead "
                    "of the library's peer-memory kernels")
    return ap.parse_args()


def f_load(x, y, z):
    return np.array(
        [np.sin(np.pi * x) * np.cos(np.pi * y), np.cos(np.pi * y) * np.sin(np.pi * z), np.sin(np.pi * z) * np.cos(np.pi * x)]
    )


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (reference algorithm, cell by cell) on a bounded sample of the same workload
def _oracle_cells(args):
    """Assemble ``cells`` tets (volume form + reference scatter) with the oracle; returns seconds."""
    n_box, lo, hi = args
    from nfempy_b200.mesh.generators import box_tet_points_cells
    from oracle import ref_assembly, ref_forms
    from oracle.ref_mesh import RefMesh
    from oracle.ref_spaces import RefProductSpace

    points, tets, _, _ = box_tet_points_cells(n_box, jitter=0.2, seed=1234)
    sub = tets[lo:hi]
    used, inv = np.unique(sub, return_inverse=True)
    mesh = RefMesh(3, points[used], [("tetra", inv.reshape(sub.shape), np.ones(len(sub), dtype=np.int64))])
    mesh.build_conformal_mesh()
    space = RefProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {})
    functions = {"rhs": f_load, "lambda": lambda x, y, z: np.ones_like(x), "mu": lambda x, y, z: np.ones_like(x)}
    t0 = time.perf_counter()
    ref_assembly.assemble(space, [("form", ref_forms.le_primal, functions, len(sub))])
    return time.perf_counter() - t0


def cpu_baseline(n_box, n_cells, workers, pool=None):
    """elements/s of the oracle port over ``workers`` processes, each assembling ``n_cells / workers`` tets."""
    per = max(1, n_cells // workers)
    jobs = [(n_box, w * per, (w + 1) * per) for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1 or pool is None:
        for job in jobs:
            _oracle_cells(job)
    else:
        pool.map(_oracle_cells, jobs)
    wall = time.perf_counter() - t0
    return per * workers / wall, per * workers, wall


def cpu_strong_baseline(space, n_cells_sample, repeats=3):
    """The "strong CPU" baseline of BASELINE.md section 3: batched numpy + scipy COO -> CSR (oracle/strong_cpu.py) on the
    first ``n_cells_sample`` tets of the SAME mesh and numbering, all host threads numpy / BLAS choose to use; best of
    ``repeats``.  Matrix only (the dominant part), lambda = mu = 1."""
    from oracle import strong_cpu

    u = space.discrete_spaces["u"]
    n = min(n_cells_sample, len(u.elements))
    dest = np.asarray(space.destination_table()[:n], dtype=np.int64)
    cell_nodes = np.asarray(u.cell_nodes()[:n], dtype=np.int64)
    points = u.mesh.points
    strong_cpu.assemble_elasticity(points, cell_nodes[:256], dest[:256], space.n_dof, 1.0, 1.0)  # warm-up (imports, einsum paths)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        strong_cpu.assemble_elasticity(points, cell_nodes, dest, space.n_dof, 1.0, 1.0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, n, best


def parity_check(torch, space, asm, plan, step, n_sample):
    """The values the timed step produced, checked in the same run: every matrix row whose cells all lie among the first
    ``n_sample`` tets is compared, entry by entry, with oracle/strong_cpu.py's assembly of those tets on the host (same
    mesh, same numbering; the strong baseline is itself pinned to the cell-by-cell oracle in the CPU tests)."""
    from oracle import strong_cpu

    u = space.discrete_spaces["u"]
    dest = np.asarray(space.destination_table(), dtype=np.int64)
    n_cells, n_loc = dest.shape
    m = min(n_sample, n_cells)
    last = np.full(space.n_dof, -1, dtype=np.int64)
    last[dest.ravel()] = np.repeat(np.arange(n_cells), n_loc)  # ascending cells: the last writer is the last cell
    rows = np.flatnonzero((last >= 0) & (last < m))
    ref_ptr, ref_idx, ref_val = strong_cpu.assemble_elasticity(u.mesh.points, np.asarray(u.cell_nodes()[:m], dtype=np.int64),
                                                               dest[:m], space.n_dof, 1.0, 1.0)
    step()
    torch.cuda.synchronize()
    indptr = plan.indptr.cpu().numpy().astype(np.int64)
    lens = indptr[rows + 1] - indptr[rows]
    if not np.array_equal(lens, ref_ptr[rows + 1].astype(np.int64) - ref_ptr[rows]):
        return {"rows_checked": int(len(rows)), "pattern_equal": False, "max_rel": None}
    pos = np.repeat(indptr[rows] - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(int(lens.sum()))
    rpos = np.repeat(ref_ptr[rows].astype(np.int64) - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(int(lens.sum()))
    dpos = torch.as_tensor(pos).to(plan.indices.device)
    got_val = asm.jacobian.values_dev.index_select(0, dpos).cpu().numpy()
    got_idx = plan.indices.index_select(0, dpos).cpu().numpy()
    want = ref_val[rpos]
    row_of = np.repeat(np.arange(len(rows)), lens)
    row_max = np.zeros(len(rows))
    np.maximum.at(row_max, row_of, np.abs(want))
    scale = np.maximum(np.abs(want), row_max[row_of])
    rel = np.abs(got_val - want) / np.where(scale > 0, scale, 1.0)
    return {"rows_checked": int(len(rows)), "values_checked": int(len(want)), "sample_cells": int(m),
            "pattern_equal": bool(np.array_equal(got_idx, ref_idx[rpos])), "max_rel": float(rel.max(initial=0.0)),
            "against": "oracle/strong_cpu.py (vectorised numpy restatement, pinned to the cell-by-cell oracle, which the "
                       "reference's own source reproduces bit for bit: tests/test_cpu_reference_pins.py)",
            "tolerance": 1e-12}


WORKLOAD = "3D Cauchy linear elasticity, vector H1 p=2, %d tets per GPU (Kuhn box n=%d, jittered), lambda=mu=1, tabulated load, CSR values + rhs"


def run_reference(args):
    """The reference's algorithm on the host cores: oracle port (cell-by-cell numpy restatement + dict-of-keys
    scatter), one worker process per core, each step a bounded sample of the same workload."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, (max(args.cpu_cells, 60 * cores) // cores) * cores)
    # the oracle builds its dict-based mesh cell by cell: every worker assembles its share of the first `per_step` tets
    # of a SMALLER box (same generator, same jitter seed, min(n, 12)^3 cubes) -- per-tet cost does not depend on the box
    n_box = min(args.n, 12)
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"  # one BLAS thread per worker, like the reference's CI (python-package.yml:17-20)
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):  # imports + first touches inside the workers
            cpu_baseline(n_box, cores * 4, cores, pool)
        times = []
        for _ in range(args.steps):
            v, cells, wall = cpu_baseline(n_box, per_step, cores, pool)
            times.append(wall)
    total = sum(times)
    value = per_step * args.steps / total
    n_full = 6 * args.n**3
    sample = ("ORACLE PORT, not nfempy itself (basix / auto_diff / petsc4py are absent here): %d tets per step = the first "
              "tets of a %d^3-cube jittered Kuhn box (same generator and seed as the %d^3 bench box, smaller so that the "
              "dict-based oracle mesh builds in seconds), split over %d worker processes, throughput extrapolated per tet.  "
              "The port's results equal the reference's own source bit for bit (tests/test_cpu_reference_pins.py), its "
              "speed does not: that source, run in the build container over stand-ins for basix / auto_diff / PETSc, did "
              "114 tets/s on one core where this port did 25 (the port differentiates by 30 residual sweeps per cell, "
              "the reference by one sweep on dual numbers) -- divide any ratio against this arm by about 4.5"
              % (per_step, n_box, args.n, cores))
    line = {
        "impl": "reference", "metric": "elements/s assembled (3D elasticity p=2)", "value": value, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % (n_full, args.n), "cells_per_gpu": n_full, "cells_timed_per_step": per_step,
                   "same_config": False, "reference_arm_is": "oracle_port"},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""

    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, window=None, probe=None):
        """``window`` / ``probe``: (t0, t1) host times of the timed region and of the identical load-probe loop that
        follows it; samples inside the timed region are used when there are at least three, else the probe's."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        def collect(lo, hi):
            sm, mx, pw, reasons = [], [], [], set()
            for stamp, ln in self.lines:
                if lo is not None and not (lo <= stamp <= hi + 0.02):
                    continue
                tok = [t.strip() for t in ln.split(",")]
                if len(tok) < 9:
                    continue
                try:
                    sm.append(float(tok[1]))
                    mx.append(float(tok[2]))
                    pw.append(float(tok[3]))
                except ValueError:
                    continue
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for name, val in zip(names, tok[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, pw, reasons

        which = "timed region"
        sm, mx, pw, reasons = collect(*window) if window else collect(None, None)
        if len(sm) < 3 and probe:
            which = "load probe (same step loop, right after the timed region)"
            sm, mx, pw, reasons = collect(*probe)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "sampled": which}


def run_graft(args):
    import torch
    import torch.distributed as dist

    from nfempy_b200 import _lib
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.mesh import box_tet_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the assembly path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # pinned host buffers are placed on the NUMA node of the allocating thread: run this rank on the cores next to its GPU
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:
            pass
    if world > 1:
        # NCCL writes its version banner / warnings to STDOUT when NCCL_DEBUG is set (it is on the GPU boxes); stdout must
        # carry the JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()

    # ---- workload: rank r owns the z-slab [r, r+1) of an n x n x (n * world) box
    # (one global mesh of world x 998 250 tets, replicated host numbering; cells partitioned into x-slabs)
    t0 = time.perf_counter()
    mesh = box_tet_mesh(args.n, jitter=0.2, seed=1234, nx=args.n * world)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {})
    t_host = time.perf_counter() - t0
    form = LEPrimalWeakForm(space)
    form.functions = {"rhs": f_load, "lambda": lambda x, y, z: np.ones_like(x), "mu": lambda x, y, z: np.ones_like(x)}
    t0 = time.perf_counter()
    sharded = None
    if world > 1:
        from nfempy_b200.assembly.sharding import ShardedAssembler

        sharded = ShardedAssembler(space, form, world, rank, device=device).prepare()
        asm, state = sharded.asm, sharded.state
        n_cells = len(sharded.shard.own_cells)
    else:
        n_cells = len(space.discrete_spaces["u"].elements)
        asm = GPUAssembler(space, None, np.zeros(space.n_dof), device=device)
        # alpha = None: the workload is the linear form at the zero state -- no K.alpha term in the residual (the API-level
        # end-to-end leg below runs with a non-zero state vector and includes it)
        asm.form_to_input_list = {"volume": ["form", range(n_cells), form, None]}
        state = asm.prepare()
    torch.cuda.synchronize()
    t_prepare = time.perf_counter() - t0
    dev, entries, datas, dests, plan, key = state
    nnz = plan.nnz
    rhs_dev = torch.zeros(plan.n_dof, dtype=torch.float64, device=device)

    def step():
        if sharded is not None:  # local fused assembly + NCCL interface-row exchange
            sharded.assemble(dist)
            return
        asm.jacobian.is_zero = True  # fresh assembly: the fused kernel overwrites every value (and every rhs row)
        asm.run(state, rhs_dev, fresh=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # dominant-kernel time for the roofline: eager launches bracketed by CUDA events on the launching stream
    for _ in range(3):
        step()
    barrier()
    lib.nfa_timing_enable(1)
    for _ in range(max(args.steps, 5)):
        step()
    barrier()
    k_ms, k_n = ctypes.c_double(), ctypes.c_int64()
    _lib.check(lib.nfa_timing_read(ctypes.byref(k_ms), ctypes.byref(k_n)), "nfa_timing_read")
    lib.nfa_timing_enable(0)
    # the timed step: on one GPU the captured CUDA graph of the same launches (GPUAssembler.capture), so that the
    # per-launch host work (Python + ctypes, ~1.3 ms per step, about what the device needs) cannot gate the device
    launches_per_step = None
    graph = None
    local_launches = 0
    p2p = False
    if sharded is not None:
        # N > 1: interface rows go straight into the owner's memory over NVLink (nfa_push_rows / nfa_pull_add_rows); the whole
        # step -- local kernels + exchange -- is then ONE CUDA graph.  Fallback (--no-p2p or no CUDA IPC): local graph + eager
        # pack / NCCL send-recv / unpack.
        p2p = (not args.no_p2p) and sharded.enable_p2p(dist)
    if sharded is not None and not args.no_graph:
        l0 = _lib.launch_count()
        sharded.local_assemble()
        local_launches = _lib.launch_count() - l0 + int(asm.last_rhs_fill)  # + the rhs fill, if one ran
        if p2p:
            local_launches += len(sharded.shard.send_rows) + len(sharded.shard.recv_rows)
            sharded.capture_step()
        else:
            sharded.capture_local()
    if sharded is None and not args.no_graph:
        l0 = _lib.launch_count()
        step()
        launches_per_step = _lib.launch_count() - l0 + int(asm.last_rhs_fill)  # + the torch fill kernel of rhs, if one ran
        graph = asm.capture(state, rhs_dev)
        eager_step = step
        step = graph.replay
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    w1 = time.time()
    ms_total = e0.elapsed_time(e1)
    if graph is not None:
        launches = launches_per_step * args.steps  # replayed kernel nodes: counted on the eager step that was captured
    else:
        launches = _lib.launch_count() - launches0 + args.steps * int(asm.last_rhs_fill)  # + the rhs fill kernel, if any
        if sharded is not None and not args.no_graph:  # the replayed kernels (+ the eager exchange kernels counted above)
            launches = _lib.launch_count() - launches0 + local_launches * args.steps
    # load probe for the clock record: the same step loop for ~0.6 s (not part of any reported time)
    p0 = time.time()
    while time.time() - p0 < 0.6:
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    p1 = time.time()
    clocks = sampler.stop(window=(w0, w1), probe=(p0, p1))
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_cells * world / (ms_step * 1e-3)

    # ---- N > 1: where the step goes, per rank (local kernels alone / whole step with the exchange), device-timed
    breakdown = None
    if sharded is not None:
        def timed(fn, sync_ranks):
            for _ in range(3):
                fn()
            if sync_ranks:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / args.steps

        step_ms_rank = timed(step, True)  # own clock of this rank around the same K steps (no max)
        dist.barrier()
        local_graph = sharded.asm.capture(sharded.state, sharded.rhs, fresh="rows")  # the local kernels alone, no exchange
        local_ms_rank = timed(local_graph.replay, False)
        dist.barrier()
        payload = int(sum(sharded._sizes[t] + int(sharded._rows_dev[t].shape[0]) for t in sharded._sizes if t[0] == "send")) * 8
        mine = {"rank": rank, "step_ms": step_ms_rank, "local_ms": local_ms_rank, "send_bytes": payload,
                "send_to": sorted(int(q) for q in sharded.shard.send_rows), "recv_from": sorted(int(q) for q in sharded.shard.recv_rows),
                "own_cells": int(n_cells), "plan_cells": int(datas[0].n_items)}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            loc = [g["local_ms"] for g in gathered]
            breakdown = {"per_rank": gathered, "local_ms_max": max(loc), "local_ms_min": min(loc),
                         "step_ms_max_over_ranks": ms_step, "exchange_and_skew_ms": ms_step - max(loc)}

    # ---- roofline of the dominant (fused) kernel; algorithmic bytes per tet, SURVEY.md 8(d)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    nnz_per_tet = nnz / n_cells
    b_kernel = 112.0 + 40.0 + 16.0 + 8.0 * nnz_per_tet  # node ids + coords gather, dof table, lambda/mu, CSR values
    rhs_folded = bool(asm._folds_load(lib, plan, entries, datas))
    if rhs_folded:
        b_kernel += 8.0 * plan.n_dof / n_cells  # the kernel also writes the right-hand-side rows ("+ RHS")
    b_step = b_kernel + 8.0 * plan.n_dof / n_cells + 27 * 3 * 8.0  # + rhs rows + tabulated load
    kernel_ms = k_ms.value / max(k_n.value, 1)
    achieved = b_kernel * n_cells / (kernel_ms * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of this exact workload
    row_items = bool(getattr(plan, "p2_stats", None))
    tpath = os.path.join(ROOT, "profiles", "r02_p2tet_traffic.json" if row_items else "r01_fused_h1_traffic.json")
    if os.path.exists(tpath) and args.n == 55 and world == 1:
        traffic = json.load(open(tpath))["traffic_bytes_per_launch"]
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "algorithmic_bytes_per_launch": b_kernel * n_cells,
        "kernel": "p2tet_elasticity_kernel" if row_items else "fused_h1_kernel<3,0>", "kernel_ms": kernel_ms, "algorithmic_bytes_per_tet": b_kernel,
        "kernel_writes_rhs": rhs_folded, "peak_source": peak_src, "step_frac": b_step * n_cells / (ms_step * 1e-3) / 1e9 / peak,
        "step_algorithmic_bytes_per_tet": b_step,
    }

    # ---- end to end through the host-buffer API: H2D of every input, kernels, D2H of CSR values + rhs
    e2e = e2e_all = None
    if not args.no_e2e:
        e2e_all = run_e2e(args, torch, dist, world, device, asm, state, form, n_cells,
                          (lambda: sharded.rhs) if sharded is not None else (lambda: rhs_dev), step, sharded)
        # N = 1: the headline end-to-end number goes through the drop-in API call itself; N > 1: the sharded step with
        # host buffers (there is no multi-process form of the reference's API to mirror)
        e2e = run_e2e_api(args, torch, device, space, form, n_cells, asm) if sharded is None else e2e_all

    varying = None
    if world == 1 and not args.no_varying:
        varying = run_varying(args, torch, lib, device, space, datas[0], n_cells, b_kernel, b_step, peak)

    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        parity = parity_check(torch, space, asm, plan, step, 40000)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cells, wall = cpu_baseline(min(args.n, 12), args.cpu_cells, 1)
        cpu = {"value": v, "unit": "elements/s", "cores": 1, "kind": "port",
               "sample": "%d tets of the same jittered Kuhn box, oracle port (cell-by-cell numpy restatement of the "
                         "reference, dict-of-keys scatter), %.1f s; the reference's own source over stand-ins is about "
                         "4.5 x faster per core than this port (DESIGN.md section 6)" % (cells, wall)}
    cpu_strong = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cells, wall = cpu_strong_baseline(space, 40000)
        cpu_strong = {"value": v, "unit": "elements/s", "cores": os.cpu_count(), "kind": "vectorised numpy + scipy COO->CSR",
                      "sample": "matrix of the first %d tets of the same mesh, best of 3, %.2f s" % (cells, wall)}
    if rank == 0:
        line = {
            "metric": "elements/s assembled (3D elasticity p=2)", "value": value, "unit": "elements/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD % (n_cells, args.n) + (
                    "" if world == 1 else "; x-slab cell partition, interface rows summed onto their owners"),
                "cells_per_gpu": n_cells, "n_dof_per_gpu": plan.n_dof, "nnz_per_gpu": nnz,
                "nonzeros_per_s": nnz * world / (ms_step * 1e-3),
                "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no explicit flush" % ((b_step * n_cells + 6 * 100 * n_cells) / 1e9),
                "plan_build_s": plan.build_seconds, "host_mesh_space_s": t_host, "prepare_s": t_prepare,
                "path": ("fused (row items)" if row_items else "fused") if plan.fused_index is not None else "two-stage",
                "launch": ("CUDA graph replay of the captured step" if graph is not None else
                           (("CUDA graph replay of the whole step: local assembly + peer-memory exchange kernels" if p2p else
                             "CUDA graph replay of the local assembly + eager NCCL exchange") if (sharded is not None and not args.no_graph)
                            else "eager launches")),
                "exchange": None if sharded is None else ("nfa_push_rows / nfa_pull_add_rows over NVLink peer memory (CUDA IPC)" if p2p
                                                          else "pack kernels + ncclSend/ncclRecv + unpack kernels"),
            },
            "roofline": roofline, "breakdown": breakdown, "varying_coefficients": varying, "parity": parity, "cpu_baseline": cpu, "cpu_strong": cpu_strong, "e2e": e2e, "e2e_all_inputs": e2e_all, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_varying(args, torch, lib, device, space, data0, n_cells, b_kernel, b_step, peak):
    """The same step with CELL-WISE VARYING Lame coefficients (the general-rho instantiation of the row-item kernel: Kf is
    applied per contribution instead of once per block -- twice the FP64 work of the constant-coefficient headline).  Same
    mesh, same load, own plan and CSR; timed like the headline (graph replay, CUDA events, args.steps steps)."""
    from nfempy_b200 import _lib
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    rng = np.random.default_rng(99)
    arrays = {"lambda": 1.0 + rng.random(n_cells), "mu": 0.5 + rng.random(n_cells),
              "rhs": None if data0.f is None else data0.f.cpu().numpy()}

    class CellwiseForm(LEPrimalWeakForm):
        def prepare(self, dev, items=None, arrays_=None):
            return LEPrimalWeakForm.prepare(self, dev, items, arrays=arrays)

    form = CellwiseForm(space)
    form.functions = {"rhs": None, "lambda": None, "mu": None}
    asm = GPUAssembler(space, None, np.zeros(space.n_dof), device=device)
    asm.form_to_input_list = {"volume": ["form", range(n_cells), form, None]}
    state = asm.prepare()
    plan = state[4]
    rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=device)

    def step():
        asm.jacobian.is_zero = True
        asm.run(state, rhs, fresh=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.nfa_timing_enable(1)
    for _ in range(max(args.steps, 5)):
        step()
    torch.cuda.synchronize()
    k_ms, k_n = ctypes.c_double(), ctypes.c_int64()
    _lib.check(lib.nfa_timing_read(ctypes.byref(k_ms), ctypes.byref(k_n)), "nfa_timing_read")
    lib.nfa_timing_enable(0)
    replay = step if args.no_graph else asm.capture(state, rhs).replay
    for _ in range(max(args.warmup, 3)):
        replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    kernel_ms = k_ms.value / max(k_n.value, 1)
    return {"coefficients": "lambda in [1, 2), mu in [0.5, 1.5) drawn per cell", "ms_per_step": ms,
            "value": n_cells / (ms * 1e-3), "unit": "elements/s", "kernel_ms": kernel_ms,
            "frac": b_kernel * n_cells / (kernel_ms * 1e-3) / 1e9 / peak,
            "step_frac": b_step * n_cells / (ms * 1e-3) / 1e9 / peak,
            "path": "fused (row items)" if getattr(plan, "p2_stats", None) else "fused"}


def run_e2e_api(args, torch, device, space, form, n_cells, asm):
    """End to end through the drop-in call, the way a Newton / time loop of the reference uses it
    (R/examples/geothermal_flow/three_fields_advection_diffusion.py:397-431): host state vector in,
    ``assembler.scatter_forms()``, host residual and ``jacobian.getValuesCSR()`` out.  Timed from the second call on (plan
    and coefficient arrays warm, as in every iteration after the first); the state vector is non-zero, so the K.alpha
    term of the residual is computed as well."""
    from nfempy_b200.assembly import CSRMatrix

    n_dof = int(space.n_dof)
    residual = np.zeros(n_dof)
    jac = CSRMatrix(n_dof)
    jac.reuse_host_buffer = True  # the loop consumes (ai, aj, av) before it reassembles
    api = type(asm)(space, jac, residual, device=device)
    api._plans = asm._plans  # the sparsity plan of the device-resident measurement (built once per space + form list)
    rng = np.random.default_rng(7)
    alpha = 1e-3 * rng.standard_normal(n_dof)
    api.form_to_input_list = {"volume": ["form", range(n_cells), form, alpha]}
    t0 = time.perf_counter()
    api.scatter_forms()
    first_call_s = time.perf_counter() - t0
    jac.getValuesCSR()
    steps = max(3, min(args.steps, 5))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        alpha += 1e-6  # the next iterate
        residual[:] = 0.0
        jac.scale(0.0)
        api.scatter_forms()
        ai, aj, av = jac.getValuesCSR()
    e1.record()
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - w0) / steps
    ms = max(e0.elapsed_time(e1) / steps, wall_ms)  # the call is synchronous: host wall clock and device time agree
    return {"value": n_cells / (ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": int(alpha.nbytes),
            "d2h_bytes_per_step": int(av.nbytes + residual.nbytes), "ms_per_step": ms, "steps": steps,
            "call": "GPUAssembler.scatter_forms() + CSRMatrix.getValuesCSR(), second call onwards",
            "first_call_s": first_call_s, "timings_last_call": api.last_timings}


def run_e2e(args, torch, dist, world, device, asm, state, form, n_cells, asm_rhs, assemble_step, sharded=None):
    """Same step, host buffers in and out: pinned coords / connectivity / coefficient / load arrays are copied to
    the device every step and the CSR values + rhs are copied back.  N > 1: a rank ships ITS share only -- the coordinates
    of the mesh nodes its cells use go up, the rows it owns (CSR segments + rhs entries, a few contiguous row blocks) come
    down; the arrays stay global-sized on the device, the PCIe traffic does not."""
    dev, entries, datas, dests, plan, key = state
    data = datas[0]
    if sharded is None:
        node_lo, node_hi = 0, int(dev.coords.shape[0])
        row_blocks = [(0, plan.n_dof)]
    else:
        used = torch.unique(data.cell_nodes)
        node_lo, node_hi = int(used.min().item()), int(used.max().item()) + 1
        row_blocks = sharded.shard.row_block_ranges()
        if len(row_blocks) > 64:  # (not the case for slab partitions: one block per entity class)
            row_blocks = [(0, plan.n_dof)]
    indptr = plan.indptr.cpu().numpy()
    val_blocks = [(int(indptr[a]), int(indptr[b])) for a, b in row_blocks]
    coords_dev = dev.coords[node_lo:node_hi]
    host_in = {
        "coords": coords_dev.cpu().pin_memory(), "cell_nodes": data.cell_nodes.cpu().pin_memory(),
        "f": data.f.cpu().pin_memory(), "lambda": data.coefs["lambda"][0].cpu().pin_memory(),
        "mu": data.coefs["mu"][0].cpu().pin_memory(),
    }
    host_vals = [torch.empty(b - a, dtype=torch.float64).pin_memory() for a, b in val_blocks]
    host_rhs = [torch.empty(b - a, dtype=torch.float64).pin_memory() for a, b in row_blocks]
    h2d = sum(t.numel() * t.element_size() for t in host_in.values())
    d2h = sum(t.numel() * 8 for t in host_vals) + sum(t.numel() * 8 for t in host_rhs)

    def step():
        coords_dev.copy_(host_in["coords"], non_blocking=True)
        data.cell_nodes.copy_(host_in["cell_nodes"], non_blocking=True)
        data.f.copy_(host_in["f"], non_blocking=True)
        data.coefs["lambda"][0].copy_(host_in["lambda"], non_blocking=True)
        data.coefs["mu"][0].copy_(host_in["mu"], non_blocking=True)
        assemble_step()  # same step as the device-resident measurement (incl. the interface exchange at N > 1)
        vals_dev, rhs_dev = asm.jacobian.values_dev, asm_rhs()
        for (a, b), h in zip(val_blocks, host_vals):
            h.copy_(vals_dev[a:b], non_blocking=True)
        for (a, b), h in zip(row_blocks, host_rhs):
            h.copy_(rhs_dev[a:b], non_blocking=True)

    steps = max(3, min(args.steps, 5))
    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"value": n_cells * world / (ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": ms, "steps": steps,
            "call": "every input array re-uploaded from pinned host memory, kernels, CSR values + rhs downloaded"
                    + ("" if sharded is None else " (per rank: its cells' inputs up, the %d row blocks it owns down)" % len(row_blocks))}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_graft(a)
