"""Scratch: host-to-host call (EnsembleSolver.solve_fixed_dt_host) for different row-block counts / streams."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import EnsembleSolver
batch, n, K = 65536, 4096, 20
s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=3, dx=3.0 / n, eps=1e-12, batch=batch)
x = torch.linspace(0, 1, s.nx, dtype=torch.float64)
host_in = torch.empty((batch, s.nx), dtype=torch.float64, pin_memory=True)
host_out = torch.empty((batch, s.nx), dtype=torch.float64, pin_memory=True)
host_in.copy_(0.5 + torch.sin(2 * np.pi * x)[None, :] * torch.rand(batch, 1, dtype=torch.float64))
dt = 0.4 * (3.0 / n) / 1.5
s.solve_fixed_dt_host(host_in, host_out, dt, 1)
torch.cuda.synchronize()
for groups, streams in ((8, 4), (16, 4), (32, 4), (64, 4), (16, 8), (32, 8), (32, 2), (64, 8), (8, 4)):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.solve_fixed_dt_host(host_in, host_out, dt, K, groups=groups, streams=streams)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"groups": groups, "streams": streams, "seconds": best, "cell_updates_per_s": batch * n * K / best}), flush=True)
