#!/usr/bin/env python
"""Kernel time of experimental builds of the row-item kernel (tools/build_variant.sh) on the bench workload, all in ONE
process: mesh, space and plan are built once with the regular library; every variant library is then loaded beside it and
its ``nfa_assemble_p2tet_elasticity`` is called on the same plan (the plan structs are plain data shared by all builds of
nfa_plan.cuh).  Each result is compared with the regular library's values.

    python tools/p2_variants.py --variants base,pf,e2 --threads 512 --plans 640:296:8,640:148:8 > gpurun_out/p2_variants.jsonl

variant spec: NAME[@THREADS[@NBUF]]; plan spec: CAP:RIFFLE:BATCH[:LANEOPT[:LANEWIN[:LOCALITY[:TYPEORDER]]]]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=55)
    ap.add_argument("--variants", default="base")
    ap.add_argument("--plans", default="640:296:8")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--varying", action="store_true", help="cell-wise lambda (general-rho instantiation)")
    ap.add_argument("--load", action="store_true", help="time nfa_assemble_p2tet_elasticity_load (matrix + right-hand side)")
    args = ap.parse_args()
    import torch

    from bench import f_load
    from nfempy_b200 import _lib
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.mesh import box_tet_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LEPrimalWeakForm

    base = _lib.load()
    mesh = box_tet_mesh(args.n, jitter=0.2, seed=1234)
    space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {})
    n_cells = len(space.discrete_spaces["u"].elements)
    form = LEPrimalWeakForm(space)
    lam = (lambda x, y, z: 1.0 + 0.5 * np.sin(3 * x) * np.ones_like(y)) if args.varying else (lambda x, y, z: np.ones_like(x))
    form.functions = {"rhs": f_load, "lambda": lam, "mu": lambda x, y, z: np.ones_like(x)}
    libs = {"base": base}

    def get_lib(name):
        if name not in libs:
            path = os.path.join(ROOT, "nfempy_b200", "csrc", "variants", "libnfa_%s.so" % name)
            lib = ctypes.CDLL(path)
            for fn in ("nfa_assemble_p2tet_elasticity", "nfa_assemble_p2tet_elasticity_load", "nfa_timing_enable",
                       "nfa_timing_read", "nfa_last_error"):
                res, at = _lib.SIGNATURES[fn]
                getattr(lib, fn).restype = res
                getattr(lib, fn).argtypes = at
            libs[name] = lib
        return libs[name]

    for pspec in args.plans.split(","):
        tok = pspec.split(":") + ["0", "48", "1", "01234"][max(0, len(pspec.split(":")) - 3):]
        cap, riffle, batch, laneopt, lanewin, local, torder = tok[:7]
        os.environ["NFA_P2TET_LOCAL"] = local
        os.environ["NFA_P2TET_TYPEORDER"] = torder
        os.environ["NFA_P2TET_CAP"], os.environ["NFA_P2TET_RIFFLE"], os.environ["NFA_P2TET_BATCH"] = cap, riffle, batch
        os.environ["NFA_P2TET_LANEOPT"], os.environ["NFA_P2TET_LANEWIN"] = laneopt, lanewin
        for k in ("NFA_P2TET_THREADS", "NFA_P2TET_NBUF"):
            os.environ.pop(k, None)
        asm = GPUAssembler(space, None, np.zeros(space.n_dof))
        asm.form_to_input_list = {"volume": ["form", range(n_cells), form, np.zeros(space.n_dof)]}
        state = asm.prepare()
        dev, entries, datas, dests, plan, key = state
        data = datas[0]
        rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=dev.device)
        asm.jacobian.is_zero = True
        asm.run(state, rhs)
        torch.cuda.synchronize()
        ref = asm.jacobian.values_dev.clone()
        ref_max = float(ref.abs().max().item())
        out = torch.empty_like(ref)
        lam_c, mu_c = data.coefs["lambda"], data.coefs["mu"]
        stream = _lib.stream_ptr()

        re4 = torch.empty((data.n_items, 10, 4), dtype=torch.float64, device=dev.device)
        form.launch_load(data.dev, data, re4, comp_pitch=4)
        rhs_ref = rhs.clone()
        rhs_out = torch.empty_like(rhs)

        def call(lib):
            if args.load:
                rc = lib.nfa_assemble_p2tet_elasticity_load(plan.bplan, _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes),
                                                            _lib.ptr(lam_c[0]), lam_c[1], _lib.ptr(mu_c[0]), mu_c[1], _lib.ptr(out),
                                                            _lib.ptr(re4), _lib.ptr(rhs_out), 1, stream)
            else:
                rc = lib.nfa_assemble_p2tet_elasticity(plan.bplan, _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), _lib.ptr(lam_c[0]),
                                                       lam_c[1], _lib.ptr(mu_c[0]), mu_c[1], _lib.ptr(out), stream)
            if rc != 0:
                raise RuntimeError((lib.nfa_last_error() or b"").decode())

        for vspec in args.variants.split(","):
            tok = vspec.split("@")
            name = tok[0]
            threads = tok[1] if len(tok) > 1 and tok[1] else None
            nbuf = tok[2] if len(tok) > 2 and tok[2] else None
            for k, v in (("NFA_P2TET_THREADS", threads), ("NFA_P2TET_NBUF", nbuf)):
                if v:
                    os.environ[k] = v
                else:
                    os.environ.pop(k, None)
            rec = {"plan": pspec, "variant": name, "threads": threads, "nbuf": nbuf, "varying": args.varying}
            try:
                lib = get_lib(name)
                out.fill_(float("nan"))
                for _ in range(3):
                    call(lib)
                torch.cuda.synchronize()
                lib.nfa_timing_enable(1)
                for _ in range(args.steps):
                    call(lib)
                torch.cuda.synchronize()
                ms, cnt = ctypes.c_double(), ctypes.c_int64()
                lib.nfa_timing_read(ctypes.byref(ms), ctypes.byref(cnt))
                lib.nfa_timing_enable(0)
                rec["kernel_ms"] = ms.value / max(cnt.value, 1)
                rec["max_rel_diff"] = float((out - ref).abs().max().item()) / ref_max
                rec["nan"] = bool(torch.isnan(out).any().item())
                if args.load:
                    rec["rhs_max_rel_diff"] = float((rhs_out - rhs_ref).abs().max().item()) / float(rhs_ref.abs().max().item())
            except Exception as exc:  # report and go on
                rec["error"] = str(exc)
            print(json.dumps(rec), flush=True)
        if getattr(plan, "p2_stats", None):
            print(json.dumps({"plan": pspec, "stats": plan.p2_stats, "plan_build_s": plan.build_seconds}), flush=True)
        del asm, state, plan, rhs, ref, out, datas, data
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
