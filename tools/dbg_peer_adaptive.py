"""Scratch: adaptive-dt history of the peer-memory slab solver vs the single-array solve."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200 import _lib as L
from pyshocks_b200.distributed import PeerRing, PeerSlabSolver
from pyshocks_b200.ensemble import EnsembleSolver

n, g = 1 << 16, 3
x = (np.arange(n) + 0.5) / n
ic = 0.5 + np.sin(2 * np.pi * x) + 0.3 * np.cos(6 * np.pi * x + 0.3)
ug = torch.from_numpy(ic).cuda()
theta, tfinal, cfl = 0.8, 0.002, 0.5 * (3.0 / n)
single = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=3.0 / n, eps=1e-12, batch=1)
u0 = torch.zeros((1, n + 2 * g), dtype=torch.float64, device="cuda"); u0[0, g:g + n] = ug
sres = single.solve_adaptive(u0, theta=theta, tfinal=tfinal, cfl_scale=cfl, check_every=1, record_dt=True)
for overlap in (False, True):
    ps = PeerSlabSolver(n_global=n, rank=0, world=1, dx=3.0 / n, edge=1024, overlap=overlap)
    ps.attach(PeerRing.local([ps.mem], 0))
    ps.load_interior(ug)
    s = ps.solver
    s.t.zero_(); s.nonfinite.zero_(); s.hp.max_abs(s.u, 1, out=s.maxabs)
    hist, mx = [], []
    while True:
        mx.append(float(s.maxabs[0]))
        L.check("c", L.lib().psk_step_control(1, theta, cfl, tfinal, L.ptr(s.maxabs), L.ptr(s.t), L.ptr(s.t), L.ptr(s.dt),
                                              L.raw_ptr(s.active), L.raw_ptr(s.nonfinite), L.stream_ptr()))
        if int(s.active.item()) == 0:
            break
        hist.append(float(s.dt[0]))
        s.maxabs.zero_()
        ps.step(s.dt, maxabs=s.maxabs)
        ps.join()
    hist = np.array(hist)
    ref = sres.dt_history[:, 0]
    k = min(len(hist), len(ref))
    bad = np.nonzero(hist[:k] != ref[:k])[0]
    print("overlap", overlap, "steps", len(hist), len(ref), "first dt mismatch", bad[:5], "max rel", np.abs(hist[:k] / ref[:k] - 1).max())
    if len(bad):
        i = bad[0]; print("  dt", hist[i], ref[i], "maxabs seen", mx[i])
    print("  state equal:", torch.equal(ps.interior(), single.u[0, g:g + n]), float((ps.interior() - single.u[0, g:g + n]).abs().max()))
    ps.ring = None; ps.solver = None; ps.mem.close()
