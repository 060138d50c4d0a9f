"""Scratch micro-benchmark run on the GPU box (not part of the product)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import EnsembleSolver

def run(batch, n, steps, math="fast", tag=""):
    h = 3.0 / n
    s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=3, dx=h, eps=1e-12, batch=batch, math=math)
    x = torch.linspace(0, 1, s.nx, device="cuda", dtype=torch.float64)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    u0 = 0.5 + torch.sin(2 * np.pi * x)[None, :] * torch.rand(batch, 1, device="cuda", dtype=torch.float64, generator=gen)
    s.load(u0)
    dt = 0.4 * h / 1.5
    s.solve_fixed_dt(None, dt, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve_fixed_dt(None, dt, steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cu = batch * n * steps / (ms * 1e-3)
    print(json.dumps({"tag": tag, "batch": batch, "n": n, "steps": steps, "ms_per_step": ms / steps, "cell_updates_per_s": cu, "hbm_frac_64B": cu * 64 / 6547.2e9}))
    return s.u.clone()

from pyshocks_b200 import _lib
a = run(65536, 4096, 20, tag="one chunk per warp")
_lib.lib().psk_set_stage_variant(5001)
b = run(65536, 4096, 20, tag="persistent + cp.async prefetch")
print("bitwise equal:", torch.equal(a, b))
_lib.lib().psk_set_stage_variant(5000)
