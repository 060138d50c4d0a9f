"""Scratch: fused exchange replayed as a CUDA graph on a one-slab ring, with the epoch state printed."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200.distributed import PeerRing, PeerSlabSolver
n = 6144
x = (np.arange(n) + 0.5) / n
ug = torch.from_numpy(0.5 + np.sin(2 * np.pi * x)).cuda()
ps = PeerSlabSolver(n_global=n, rank=0, world=1, dx=3.0 / n, timeout_s=1.0)
ps.attach(PeerRing.local([ps.mem], 0))
ps.load_interior(ug)
dtt = torch.full((1,), 1e-5, dtype=torch.float64, device="cuda")
def show(tag):
    torch.cuda.synchronize()
    print(tag, "host epoch", ps.epoch, "flags", ps.mem.flags.tolist(), "epoch_dev", ps.epoch_dev.tolist(), "timed_out", int(ps.timed_out), flush=True)
show("loaded")
ps.solve_fixed_dt(dtt, 2); show("2 plain steps")
ps.solve_fixed_dt(dtt, 6, graph=True); show("6 steps graph")
ps.solve_fixed_dt(dtt, 7, graph=True); show("7 steps graph")
