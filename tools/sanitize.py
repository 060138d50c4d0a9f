"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver
from pyshocks_b200.path import HotPath
for flux in ("rusanov", "lf"):
    B, n, g = 3, 300, 3
    kw = dict(equation="burgers", flux=flux, rec="wenojs53", bc="periodic", n=n, g=g, dx=3.0 / n, eps=1e-12)
    u0 = torch.from_numpy(0.5 + np.sin(np.linspace(0, 6.28, n + 2 * g))[None, :].repeat(B, 0)).cuda()
    for math in ("fast", "strict"):
        s = EnsembleSolver(batch=B, math=math, **kw)
        s.solve_fixed_dt(u0, 1e-3, 2)
        s.solve_adaptive(u0, theta=0.9, tfinal=0.004, cfl_scale=0.5 * 3.0 / n, check_every=1)
    adj = AdjointEnsemble(EnsembleSolver(batch=B, **kw), nsteps=3, dt=1e-3, segment=2)
    adj.gradient_half_l2(u0)
    hp = HotPath(math="fast", **kw)
    x = u0[0].contiguous()
    hp.apply_operator(x); hp.numerical_flux(hp.apply_boundary(x)); hp.reconstruct(x); hp.apply_operator_vjp(x, x.clone())
    hp.ssprk33_step(x, torch.tensor([1e-3], dtype=torch.float64, device="cuda"), ghost_rows=True)
    hp.solve_rows(x.clone(), tfinal=0.004, theta=0.9, cfl_scale=0.5 * 3.0 / n, max_steps=50, tape=True, record_dt=True)
# ESWENO32 (reconstruction + Burgers scheme), FAST and STRICT
for math in ("fast", "strict"):
    n, g = 300, 2
    hp = HotPath(equation="burgers", flux="esweno32", rec="esweno32", bc="periodic", n=n, g=g, dx=3.0 / n, eps=2.7e-4,
                 delta=1e-4, math=math)
    x = torch.from_numpy(0.5 + np.sin(np.linspace(0, 6.28, n + 2 * g))).cuda()
    hp.apply_operator(x); hp.numerical_flux(hp.apply_boundary(x)); hp.reconstruct(x)
    hp.ssprk33_step(x, torch.tensor([1e-3], dtype=torch.float64, device="cuda"), ghost_rows=True)
# slab decomposition over peer memory (all three forms, rings held by this process) + adaptive dt
from pyshocks_b200.distributed import PeerRing, PeerSlabSolver
n = 6144
ug = torch.from_numpy(0.5 + np.sin(np.linspace(0, 6.28, n))).cuda()
dtt = torch.full((1,), 1e-4, dtype=torch.float64, device="cuda")
for mode in ("serial", "overlap", "fused"):
    slabs = [PeerSlabSolver(n_global=n, rank=r, world=2, dx=3.0 / n, edge=256, overlap=(mode == "overlap"),
                            fused=(mode == "fused"), timeout_s=30.0) for r in range(2)]
    for r, sl in enumerate(slabs):
        sl.attach(PeerRing.local([t.mem for t in slabs], r))
    for sl in slabs:
        sl.load_interior(ug[sl.first : sl.first + sl.n_local])
    for _ in range(2):
        for stage in (1, 2, 3):
            for sl in slabs:
                sl.run_stage(stage, dtt)
    for sl in slabs:
        sl.join(); sl.check(); sl.ring = None; sl.solver = None; sl.mem.close()
one = PeerSlabSolver(n_global=n, rank=0, world=1, dx=3.0 / n, timeout_s=30.0)
one.attach(PeerRing.local([one.mem], 0))
one.load_interior(ug)
one.solve_adaptive(theta=0.9, tfinal=0.0005, cfl_scale=0.5 * 3.0 / n)
one.ring = None; one.solver = None; one.mem.close()
torch.cuda.synchronize()
print("sanitize target done")
