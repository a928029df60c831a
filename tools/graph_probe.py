import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nfempy_b200.assembly import GPUAssembler
from nfempy_b200.mesh import box_tet_mesh
from nfempy_b200.spaces import ProductSpace
from nfempy_b200.weak_forms import LEPrimalWeakForm
from nfempy_b200 import _lib
import bench
mesh = box_tet_mesh(55, jitter=0.2, seed=1234)
space = ProductSpace({"u": (3, 3, "Lagrange", 2, mesh)}); space.build_structures({"u": [1]}, {})
form = LEPrimalWeakForm(space)
form.functions = {"rhs": bench.f_load, "lambda": lambda x, y, z: np.ones_like(x), "mu": lambda x, y, z: np.ones_like(x)}
n = len(space.discrete_spaces["u"].elements)
asm = GPUAssembler(space, None, np.zeros(space.n_dof))
asm.form_to_input_list = {"volume": ["form", range(n), form, np.zeros(space.n_dof)]}
state = asm.prepare()
plan = state[4]
rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device="cuda")
def step():
    asm.jacobian.is_zero = True
    rhs.zero_()
    asm.run(state, rhs)
for _ in range(3): step()
torch.cuda.synchronize()
def timeit(fn, k=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("eager ms/step", timeit(step))
t0=time.perf_counter()
for _ in range(20): step()
print("host enqueue ms/step", (time.perf_counter()-t0)/20*1e3); torch.cuda.synchronize()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        step()
    print("graph ms/step", timeit(g.replay))
except Exception as ex:
    print("graph capture failed:", repr(ex)[:300])
