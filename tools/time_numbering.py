#!/usr/bin/env python
"""Host time of the mesh + space numbering with the device-accelerated primitives on and off:
    python tools/time_numbering.py --n 55 --nx 55,220,440
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=55)
    ap.add_argument("--nx", default="55")
    ap.add_argument("--numpy-too", default="55", help="comma list of nx values also timed with the numpy path")
    args = ap.parse_args()
    import numpy as np
    import torch

    from nfempy_b200.assembly.sharding import ShardPlan
    from nfempy_b200.mesh import _accel, box_tet_mesh
    from nfempy_b200.spaces import ProductSpace

    torch.zeros(1, device="cuda")
    numpy_too = {int(v) for v in args.numpy_too.split(",") if v}
    for nx in [int(v) for v in args.nx.split(",")]:
        for accel in ([True, False] if nx in numpy_too else [True]):
            _accel.enable(accel)
            t0 = time.perf_counter()
            mesh = box_tet_mesh(args.n, jitter=0.2, seed=1234, nx=nx)
            t1 = time.perf_counter()
            space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
            space.build_structures({"u": [1]}, {})
            dest = space.destination_table()
            t2 = time.perf_counter()
            world = max(1, nx // args.n)
            ShardPlan(dest, space.n_dof, world, 0)
            t3 = time.perf_counter()
            print(json.dumps({"n": args.n, "nx": nx, "tets": int(dest.shape[0]), "device_primitives": accel, "mesh_s": t1 - t0,
                              "space_s": t2 - t1, "shard_plan_s": t3 - t2, "world": world}), flush=True)
            del mesh, space, dest


if __name__ == "__main__":
    main()
