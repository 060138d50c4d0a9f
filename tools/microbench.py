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
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_fixed_dt(None, dt, steps); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cu = batch * n * steps / (best * 1e-3)
    print(json.dumps({"tag": tag, "batch": batch, "n": n, "ms_per_step": best / steps, "cell_updates_per_s": cu, "hbm_frac_64B": cu * 64 / 6547.2e9}))

from pyshocks_b200 import _lib
for wmax in (8, 5, 4, 8):
    _lib.lib().psk_set_stage_variant(4000 + wmax)
    run(65536, 4096, 20, tag=f"max warps per CTA {wmax}")
_lib.lib().psk_set_stage_variant(4008)
