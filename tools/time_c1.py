import sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tests/golden")
import pyshocks_b200 as ps
from pyshocks_b200 import burgers, timestepping
from pyshocks_b200.reconstruction import make_reconstruction_from_name
from pyshocks_b200.scalar import PeriodicBoundary
from common import load_golden
S = load_golden("solve_c1")
scheme = burgers.make_scheme_from_name("rusanov", rec=make_reconstruction_from_name("wenojs53"))
grid = ps.make_uniform_cell_grid(a=-1.5, b=1.5, n=256, nghosts=3); bc = PeriodicBoundary()
u0 = torch.from_numpy(S["rusanov_u0"]).cuda()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = timestepping.solve(scheme, grid, bc, u0, tfinal=1.0); torch.cuda.synchronize()
    print("single launch solve:", (time.perf_counter() - t0) * 1e3, "ms for", int(res["iteration"][0]), "steps")
stepper = timestepping.SSPRK33(predict_timestep=lambda t_, u_: ps.predict_timestep(scheme, grid, bc, t_, u_), source=ps.bind_operator(scheme, grid, bc), checkpoint=None)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for ev in timestepping.step(stepper, u0, tfinal=1.0): pass
    torch.cuda.synchronize()
    print("step-by-step API:", (time.perf_counter() - t0) * 1e3, "ms")
from oracle.c_oracle import COracle
from oracle import pyshocks_oracle as po
g = po.make_grid(-1.5, 1.5, 256, 3)
co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=256, g=3, batch=1, dx=g.h, eps=1e-12)
t0 = time.perf_counter(); co.solve_adaptive(S["rusanov_u0"], 1.0, 0.5 * g.h, 1.0); print("C oracle (1 core):", (time.perf_counter() - t0) * 1e3, "ms")
sch = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53"))
t0 = time.perf_counter()
for _ in po.step(lambda t, u: po.apply_operator(sch, g, po.Periodic(), t, u), lambda t, u: po.predict_timestep(sch, g, u), S["rusanov_u0"], tfinal=1.0): pass
print("NumPy oracle:", (time.perf_counter() - t0) * 1e3, "ms")
