#!/usr/bin/env python
"""Sweep of the row-item kernel's launch parameters on the bench workload (config 2, 998 250 tets): for every chunk cap
(one plan build each) x kernel variant (staged / global-load) x threads x buffers, the kernel time by CUDA events inside
the library (nfa_timing_*) and a bit-for-bit comparison of the CSR values against the global-load kernel's.

    python tools/p2_sweep.py --caps 320,384,448 --threads 512,640,768 --nbuf 0 > gpurun_out/p2_sweep.jsonl
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=55)
    ap.add_argument("--caps", default="384")
    ap.add_argument("--threads", default="512,640,768")
    ap.add_argument("--nbuf", default="0", help="comma list; 0 = as many as fit")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--env", default="", help="semicolon list of extra environment settings to cross, e.g. NFA_P2TET_REC_TMA=0;NFA_P2TET_REC_TMA=1")
    ap.add_argument("--varying", action="store_true", help="cell-wise lambda/mu (the general-rho instantiation)")
    args = ap.parse_args()
    import torch

    from bench import f_load
    from nfempy_b200 import _lib
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.mesh import box_tet_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    lib = _lib.load()
    mesh = box_tet_mesh(args.n, jitter=0.2, seed=1234)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {})
    n_cells = len(space.discrete_spaces["u"].elements)
    form = LEPrimalWeakForm(space)
    if args.varying:
        lam = lambda x, y, z: 1.0 + 0.5 * np.sin(3 * x) * np.ones_like(y)  # noqa: E731
    else:
        lam = lambda x, y, z: np.ones_like(x)  # noqa: E731
    form.functions = {"rhs": f_load, "lambda": lam, "mu": lambda x, y, z: np.ones_like(x)}

    def kernel_ms(asm, state, rhs):
        def step():
            asm.jacobian.is_zero = True
            rhs.zero_()
            asm.run(state, rhs)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        lib.nfa_timing_enable(1)
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        _lib.check(lib.nfa_timing_read(ctypes.byref(ms), ctypes.byref(cnt)), "nfa_timing_read")
        lib.nfa_timing_enable(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return ms.value / max(cnt.value, 1), e0.elapsed_time(e1) / args.steps

    for cap in [int(c) for c in args.caps.split(",")]:
        os.environ["NFA_P2TET_CAP"] = str(cap)
        t0 = time.perf_counter()
        asm = GPUAssembler(space, None, np.zeros(space.n_dof))
        asm.form_to_input_list = {"volume": ["form", range(n_cells), form, np.zeros(space.n_dof)]}
        state = asm.prepare()
        plan = state[4]
        rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=state[0].device)
        torch.cuda.synchronize()
        t_prep = time.perf_counter() - t0
        os.environ["NFA_P2TET_STAGE"] = "0"
        os.environ.pop("NFA_P2TET_THREADS", None)
        os.environ.pop("NFA_P2TET_NBUF", None)
        k_ms, s_ms = kernel_ms(asm, state, rhs)
        ref_vals = asm.jacobian.values_dev.clone()
        print(json.dumps({"cap": cap, "stage": 0, "threads": 512, "nbuf": 0, "kernel_ms": k_ms, "step_ms": s_ms,
                          "prepare_s": t_prep, "stats": plan.p2_stats}), flush=True)
        os.environ["NFA_P2TET_STAGE"] = "1"
        for th, nb, extra in [(int(t), int(b), x) for t in args.threads.split(",") for b in args.nbuf.split(",")
                              for x in (args.env.split(";") if args.env else [""])]:
            if True:
                if extra:
                    os.environ[extra.split("=")[0]] = extra.split("=")[1]
                os.environ["NFA_P2TET_THREADS"] = str(th)
                if nb:
                    os.environ["NFA_P2TET_NBUF"] = str(nb)
                else:
                    os.environ.pop("NFA_P2TET_NBUF", None)
                try:
                    k_ms, s_ms = kernel_ms(asm, state, rhs)
                    d = (asm.jacobian.values_dev - ref_vals).abs().max().item() / ref_vals.abs().max().item()
                    same = d  # max |difference| relative to max |value| (instantiations may contract FMAs differently)
                    err = None
                except Exception as exc:  # report and go on with the next variant
                    k_ms = s_ms = None
                    same, err = False, str(exc)
                print(json.dumps({"cap": cap, "stage": 1, "threads": th, "nbuf": nb, "env": extra, "kernel_ms": k_ms,
                                  "step_ms": s_ms, "max_rel_diff_to_global_load_kernel": same, "error": err}), flush=True)
        del asm, state, plan, rhs, ref_vals
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
