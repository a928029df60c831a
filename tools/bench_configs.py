#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs other than the headline one (which bench.py covers):

  config 1  2-D Poisson, H1 p=1, 64 x 64 x 2 triangles            (launch/latency bound; time only)
  config 2' 3-D elasticity p=2 through the TWO-STAGE path           (for comparison with the fused path)
  config 2" 3-D elasticity p=2 with the example driver's scatter policy (filtered_upper + explicit diagonal)
  config 3  3-D mixed Poisson, RT_k x P_{k-1}^disc, k = 1, 2, 3     (k = 1 by row owner, csrc/nfa_hdivrows.cu; k = 2, 3
            two-stage; k = 3 on a smaller box because the generic plan indexes local entries with 31 bits); the
            degenerate-elliptic form at k = 1 (row owner) and k = 2 (two-stage)
  config 4  mixed Cosserat elasticity at the reference example's orders (two-stage; tools/bench_config4.py runs the
            BASELINE order through the streamed assembler)
  config 5  two-flux transport, 100 reassemblies on one plan: one box, and the mixed-dimensional fractured box (row owner;
            tools/bench_config5_sharded.py runs it on N GPUs)

One JSON line per case: elements/s with inputs resident in HBM (CUDA events, K steps after W warm-ups), nnz, local
sizes, plan-build seconds, and the oracle's rate on a few cells of the same mesh (1 core) for scale.
Run on the GPU box:  python tools/bench_configs.py > gpurun_out/configs.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_steps(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def graph_ms(torch, asm, state, rhs, steps):
    """The same reassembly replayed from one CUDA graph (``GPUAssembler.capture``): no per-launch host work."""
    try:
        graph = asm.capture(state, rhs)
    except Exception as exc:  # value-dependent patterns cannot be captured
        print("graph capture failed: %r" % (exc,), file=sys.stderr)
        return None
    return time_steps(torch, graph.replay, steps, 3)


def oracle_rate(case, n_cells):
    """elements/s of the oracle on the first ``n_cells`` cells of the case's volume form (1 core)."""
    from oracle import ref_assembly

    space = case.ref_space()
    kind, form, functions, total = case.ref_forms_fn(space)[0]
    n_cells = min(n_cells, total)
    n = space.n_dof
    res, jac = np.zeros(n), ref_assembly.DokMatrix(n)
    t0 = time.perf_counter()
    for i in range(n_cells):
        ref_assembly.scatter_form(space, form, functions, i, res, jac, np.zeros(n))
    return n_cells / (time.perf_counter() - t0)


ONLY = ""


def wanted(name):
    """--only takes comma-separated pieces of case names"""
    return not ONLY or any(tok in name for tok in ONLY.split(","))


def path_of(plan):
    if plan.fused_index is not None:
        return "fused"
    return "row owner (records + rows, no local matrices)" if plan.rows else "two-stage"


def rows_roofline(state, ms):
    """HBM roofline of an assembly on the row-owner path: ALGORITHMIC bytes = every per-item input array of the direct
    forms (connectivity, orientation cases, coefficient / load arrays as stored, destination table), the coordinates, the
    state vectors, and the outputs (8 B per stored CSR value, 8 B per right-hand-side row) -- records and plan tables are
    not counted -- over the measured time of the whole assembly, against the measured copy bandwidth."""
    import torch

    dev, entries, datas, dests, plan, _ = state
    if not plan.rows:
        return None

    def tensors(obj, n, seen):
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda and obj.ndim >= 1 and obj.shape[0] == n and obj.data_ptr() not in seen:
                seen.add(obj.data_ptr())
                return obj.numel() * obj.element_size()
            return 0
        if isinstance(obj, dict):
            return sum(tensors(v, n, seen) for v in obj.values())
        if isinstance(obj, (tuple, list)):
            return sum(tensors(v, n, seen) for v in obj)
        return 0

    total, seen = 0, set()
    for i in plan.rows:
        data = datas[i]
        total += tensors({k: v for k, v in data.__dict__.items() if k not in ("x_host", "dev")}, data.n_items, seen)
        total += dests[i].numel() * dests[i].element_size()
        total += data.dev.coords.numel() * 8
        states = data.alpha_dev if isinstance(data.alpha_dev, tuple) else (data.alpha_dev,)
        total += sum(a.numel() * 8 for a in states if a is not None)
    total += int(plan.nnz) * 8 + int(plan.n_dof) * 8
    peak, source = 6548.2, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, source = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    achieved = total / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "algorithmic_bytes_per_assembly": total, "kernels": "rt1_records_kernel + rows_expand_kernel (+ boundary forms)",
            "peak_source": source}


def run_case(name, case, small_case, steps, warmup, allow_fused=True, oracle_cells=40):
    import torch

    if not wanted(name):
        return
    from nfempy_b200 import _lib
    from nfempy_b200.assembly import GPUAssembler

    t0 = time.perf_counter()
    space = case.space()
    t_space = time.perf_counter() - t0
    asm = GPUAssembler(space, None, np.zeros(space.n_dof), allow_fused=allow_fused)
    asm.form_to_input_list = case.gpu_forms(space, None)
    t0 = time.perf_counter()
    state = asm.prepare()
    torch.cuda.synchronize()
    t_prepare = time.perf_counter() - t0
    plan = state[4]
    rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=state[0].device)

    def step():
        asm.jacobian.is_zero = True
        rhs.zero_()
        asm.run(state, rhs)

    l0 = _lib.launch_count()
    ms = time_steps(torch, step, steps, warmup)
    launches = (_lib.launch_count() - l0) / (steps + warmup)
    first = next(iter(space.discrete_spaces.values()))
    n_cells = len(first.elements)
    n_loc = int(space.destination_table().shape[1])
    line = {
        "case": name, "cells": n_cells, "n_dof": int(space.n_dof), "nnz": int(plan.nnz), "n_loc": n_loc,
        "path": path_of(plan), "ms_per_step": ms,
        "elements_per_s": n_cells / (ms * 1e-3), "nonzeros_per_s": plan.nnz / (ms * 1e-3),
        "launches_per_step": launches, "plan_build_s": plan.build_seconds, "host_space_s": t_space,
        "prepare_s": t_prepare, "steps": steps, "warmup": warmup, "dtype": "f64",
        "local_bytes_per_cell_two_stage": 8 * (n_loc * n_loc + n_loc),
    }
    if plan.rows:
        line["roofline"] = rows_roofline(state, ms)
    if small_case is not None:
        line["oracle_elements_per_s_1core"] = oracle_rate(small_case, oracle_cells)
    print(json.dumps(line), flush=True)
    del asm, state, plan
    torch.cuda.empty_cache()


def run_filtered(mesh, steps, warmup):
    """Config 2 with the scatter policy the reference's example driver actually uses
    (R/examples/cauchy_elasticity/cg_elasticity_example_1.py:92-145): value-filtered entries, upper triangle ("sbaij"),
    explicit diagonal.  The pattern depends on the values, so this takes the two-stage path: local matrices, one pass for
    the masks, a plan over the kept entries (built at the first assembly, reused while the masks do not change), reduce."""
    import torch

    if not wanted("config2 filtered_upper example-driver policy"):
        return
    from nfempy_b200.assembly import GPUAssembler
    from tests.parity_helpers import elasticity_case

    case = elasticity_case(mesh=mesh, with_bc=False)
    space = case.space()
    asm = GPUAssembler(space, None, np.zeros(space.n_dof), pattern_policy="filtered_upper", explicit_diagonal=True)
    asm.form_to_input_list = case.gpu_forms(space, None)
    t0 = time.perf_counter()
    asm.scatter_forms()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    state = asm.prepare()
    rhs = torch.zeros(int(space.n_dof), dtype=torch.float64, device=state[0].device)

    def step():
        asm.jacobian.scale(0.0)
        asm.run(state, rhs, fresh=True)

    ms = time_steps(torch, step, steps, warmup)
    n_cells = len(space.discrete_spaces["u"].elements)
    plan = asm._filtered_plan
    print(json.dumps({
        "case": "config2, example-driver scatter policy (filtered_upper + explicit diagonal), two-stage path",
        "cells": n_cells, "n_dof": int(space.n_dof), "nnz": int(plan.nnz), "n_loc": 30, "path": "two-stage (value-dependent pattern)",
        "first_assembly_s": first, "plan_build_s": plan.build_seconds, "plan_reused": int(plan.reused),
        "ms_per_step": ms, "elements_per_s": n_cells / (ms * 1e-3), "steps": steps, "warmup": warmup, "dtype": "f64",
    }), flush=True)
    del asm, state, plan
    torch.cuda.empty_cache()


def run_sundus(mesh, n_steps):
    import torch

    if not wanted("config5 sundus"):
        return
    from nfempy_b200.assembly import GPUAssembler
    from tests.parity_helpers import SundusCase

    case = SundusCase(dim=3, k=1, mesh=mesh)
    space = case.space()
    a_n, a = case.alphas(space.n_dof)
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    asm.form_to_input_list = case.gpu_forms(space, (a_n, a))
    t0 = time.perf_counter()
    asm.scatter_forms()  # first assembly: callables, uploads, plan build, kernels, residual download
    first = time.perf_counter() - t0
    state = asm.prepare()
    plan = state[4]
    rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=state[0].device)

    def step():
        asm.jacobian.is_zero = True
        rhs.zero_()
        asm.run(state, rhs)

    ms = time_steps(torch, step, n_steps, 3)
    ms_graph = graph_ms(torch, asm, state, rhs, n_steps)
    # a full API step of a time loop: new time -> loads re-evaluated on the host, kernels, residual downloaded
    t0 = time.perf_counter()
    for k in range(3):
        asm.form_to_input_list = case.gpu_forms(space, (a_n, a))
        asm.jacobian.scale(0.0)
        asm.residual[:] = 0.0
        asm.scatter_forms()
    api = (time.perf_counter() - t0) / 3
    n_cells = len(space.discrete_spaces["mp"].elements)
    print(json.dumps({
        "case": "config5: two-flux transport (Sundus) RT1 x RT1 x P0 x P0, implicit Euler + reaction, %d reassemblies "
                "on one plan" % n_steps,
        "cells": n_cells, "n_dof": int(space.n_dof), "nnz": int(plan.nnz), "n_loc": 10, "path": path_of(plan),
        "roofline": rows_roofline(state, ms),
        "ms_per_reassembly_kernels": ms, "ms_per_reassembly_graph_replay": ms_graph,
        "elements_per_s": n_cells / (ms * 1e-3), "plans_built": len(asm._plans),
        "first_assembly_s": first, "api_step_s_with_host_callables": api, "plan_build_s": plan.build_seconds,
        "steps": n_steps, "dtype": "f64",
    }), flush=True)


def run_sundus_md(n, n_steps):
    """BASELINE config 5 as stated: the two-flux transport model on a MIXED-DIMENSIONAL fractured 3-D mesh -- Sundus forms in
    the box and on the fracture plane, boundary forms, Robin coupling of both fluxes across the fracture
    (tests/md_helpers.py), one stacked system, plan built once, ``n_steps`` reassemblies."""
    import torch

    if not wanted("config5 md sundus"):
        return
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.mesh import fractured_box_tet_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import interface_triples
    from tests.md_helpers import md_sundus_forms, md_sundus_spaces

    t0 = time.perf_counter()
    mesh, interface = fractured_box_tet_mesh(n, jitter=0.15)
    spaces = md_sundus_spaces(mesh, ProductSpace)
    triples = interface_triples(spaces[0], spaces[1], interface, "mp")
    mp0 = spaces[0].discrete_spaces["mp"]
    outer = np.flatnonzero(mesh.material[mp0.bc_element_ids] != 50)
    t_host = time.perf_counter() - t0
    n_dof = spaces[1].dof_shift + spaces[1].n_dof
    a_n, a = 0.3 * np.cos(np.arange(n_dof) * 0.37) + 0.05, 0.2 * np.sin(np.arange(n_dof) * 0.11)
    asm = GPUAssembler(spaces, None, np.zeros(n_dof))
    forms = md_sundus_forms(spaces, triples, outer, (a_n, a), 0.0)
    asm.form_to_input_list = forms
    t0 = time.perf_counter()
    asm.scatter_forms()
    first = time.perf_counter() - t0
    state = asm.prepare()
    plan = state[4]
    rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=state[0].device)

    def step():
        asm.jacobian.is_zero = True
        rhs.zero_()
        asm.run(state, rhs)

    ms = time_steps(torch, step, n_steps, 3)
    ms_graph = graph_ms(torch, asm, state, rhs, n_steps)
    # the time loop through the API: a new time per step (loads re-evaluated on the host from the cached points), a new
    # iterate, kernels, residual downloaded
    t0 = time.perf_counter()
    for k in range(5):
        a_n += 1e-3
        for key, entry in forms.items():
            if "time_dependent" in entry[0]:
                entry[4] = 0.1 * (k + 1)
        asm.jacobian.scale(0.0)
        asm.residual[:] = 0.0
        asm.scatter_forms()
    api = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for k in range(5):  # Newton iterations inside one time step: t fixed, only the iterate changes
        a_n += 1e-3
        asm.jacobian.scale(0.0)
        asm.residual[:] = 0.0
        asm.scatter_forms()
    api_newton = (time.perf_counter() - t0) / 5
    split = dict(asm.last_timings)
    n_cells = len(mp0.elements)
    print(json.dumps({
        "case": "config5 (mixed-dimensional): two-flux transport RT1 x RT1 x P0 x P0 in the fractured box (z = 1/2) + the same "
                "fields on the fracture + Robin coupling of both fluxes, %d reassemblies on one plan" % n_steps,
        "cells": n_cells, "fracture_cells": int(len(triples)), "n_dof": int(n_dof), "nnz": int(plan.nnz), "path": path_of(plan),
        "roofline": rows_roofline(state, ms),
        "ms_per_reassembly_kernels": ms, "ms_per_reassembly_graph_replay": ms_graph,
        "elements_per_s": (n_cells + len(triples)) / (ms * 1e-3), "plans_built": len(asm._plans),
        "first_assembly_s": first, "api_step_s_new_time": api, "api_step_s_newton_iteration": api_newton, "api_newton_split_s": split,
        "plan_build_s": plan.build_seconds, "host_mesh_space_s": t_host, "steps": n_steps, "dtype": "f64",
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="small meshes (smoke run)")
    ap.add_argument("--only", default="", help="run only the cases whose name contains this text")
    args = ap.parse_args()
    global ONLY
    ONLY = args.only
    from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
    from tests.parity_helpers import degenerate_case, elasticity_case, laplace_case, mixed_poisson_case, tensor_mixed_case

    q = args.quick
    # config 1
    run_case("config1: 2D Poisson H1 p=1, 64x64x2 triangles + Dirichlet penalty",
             laplace_case(dim=2, k=1, mesh=square_tri_mesh(16 if q else 64)), laplace_case(n=4, dim=2, k=1),
             args.steps, args.warmup)
    # config 2 through the two-stage path (same mesh as bench.py unless --quick)
    mesh2 = box_tet_mesh(8 if q else 55, jitter=0.2)
    run_case("config2 (two-stage path): 3D elasticity H1 p=2", elasticity_case(mesh=mesh2, with_bc=False),
             elasticity_case(n=2, with_bc=False), args.steps, args.warmup, allow_fused=False)
    run_case("config2 (fused path) + Dirichlet penalty on all six sides", elasticity_case(mesh=mesh2, with_bc=True),
             None, args.steps, args.warmup)
    run_filtered(mesh2, args.steps, args.warmup)
    del mesh2
    # config 3
    mesh3 = box_tet_mesh(8 if q else 69, jitter=0.2)
    for k in (1, 2):
        run_case("config3: 3D mixed Poisson RT%d x P%d disc + natural BC" % (k, k - 1),
                 mixed_poisson_case(dim=3, k=k, mesh=mesh3), mixed_poisson_case(dim=3, k=k, n=2), args.steps,
                 args.warmup)
    run_case("config3: DegenerateEllipticWeakForm RT1 x P0 disc + BC", degenerate_case("degenerate", 3, 1, mesh=mesh3),
             degenerate_case("degenerate", 3, 1, n=2), args.steps, args.warmup, oracle_cells=20)
    # config 5: reassembly of the two-flux transport system, plan built once, 100 "time steps" of kernels
    run_sundus(mesh3, 20 if q else 100)
    del mesh3
    run_sundus_md(8 if q else 68, 20 if q else 100)
    mesh3s = box_tet_mesh(6 if q else 40, jitter=0.2)
    run_case("config3: 3D mixed Poisson RT3 x P2 disc + natural BC (40^3 x 6 tets: 31-bit local-entry index)",
             mixed_poisson_case(dim=3, k=3, mesh=mesh3s), mixed_poisson_case(dim=3, k=3, n=1), args.steps, args.warmup,
             oracle_cells=6)
    # config 3, other form classes of the same spaces
    mesh3d = box_tet_mesh(6 if q else 48, jitter=0.2)
    run_case("config3: DegenerateEllipticWeakForm RT2 x P1 disc + BC", degenerate_case("degenerate", 3, 2, mesh=mesh3d),
             degenerate_case("degenerate", 3, 2, n=2), args.steps, args.warmup, oracle_cells=20)
    del mesh3d
    # config 4: mixed Cosserat elasticity (scaled), the reference example's two methods.  Explicit local matrices:
    # n_loc^2 * 12 B per cell bounds the mesh that fits (54 dofs: 35 kB / cell, 159 dofs: 303 kB / cell).
    mesh4 = box_tet_mesh(4 if q else 30, jitter=0.2)
    run_case("config4: LCEScaledDualWeakForm BDM1 x RT1 x P0 x P0 (k_order 0, 54 local dofs) + BC",
             tensor_mixed_case("lce_scaled", 3, "BDM", "RT", 1, mesh=mesh4), tensor_mixed_case("lce_scaled", 3, n=1),
             args.steps, args.warmup, oracle_cells=12)
    run_case("config4: LCEScaledRieszMapWeakForm BDM1 x RT1 x P0 x P0 (preconditioner matrix)",
             tensor_mixed_case("lce_scaled_riesz", 3, "BDM", "RT", 1, mesh=mesh4),
             tensor_mixed_case("lce_scaled_riesz", 3, n=1), args.steps, args.warmup, oracle_cells=12)
    del mesh4
    mesh4b = box_tet_mesh(3 if q else 14, jitter=0.2)
    run_case("config4: LCEScaledDualWeakForm BDM2 x RT2 x P1 x P1 (k_order 1, 159 local dofs) + BC",
             tensor_mixed_case("lce_scaled", 3, "BDM", "RT", 2, mesh=mesh4b),
             tensor_mixed_case("lce_scaled", 3, "BDM", "RT", 2, n=1), args.steps, args.warmup, oracle_cells=6)


if __name__ == "__main__":
    main()
