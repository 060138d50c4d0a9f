"""Profiling target (scratch): a few fused SSPRK33 steps on a mid-size ensemble."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import EnsembleSolver
batch, n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 4096
h = 3.0 / n
s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=3, dx=h, eps=1e-12, batch=batch)
x = torch.linspace(0, 1, s.nx, device="cuda", dtype=torch.float64)
s.load(0.5 + torch.sin(2 * np.pi * x)[None, :] * torch.rand(batch, 1, device="cuda", dtype=torch.float64))
s.solve_fixed_dt(None, 0.4 * h / 1.5, 4)
torch.cuda.synchronize()
