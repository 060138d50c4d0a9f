import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from pyshocks_b200 import _lib
from pyshocks_b200.ensemble import AdjointEnsemble
from test_gpu_properties import _solver, _ensemble_ic, G

def check(B, n, nsteps, variant, smooth_v):
    _lib.lib().psk_set_adjoint_variant(variant)
    s = _solver(B, n)
    u0 = _ensemble_ic(B, n, seed=20261018)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    gen = torch.Generator(device="cuda").manual_seed(1)
    v = torch.zeros_like(u0); w = torch.zeros_like(u0)
    if smooth_v:
        xh = ((torch.arange(n, device="cuda", dtype=torch.float64) + 0.5) / n)[None, :]
        v[:, G:G+n] = torch.cos(2*np.pi*3*xh + 0.3); w[:, G:G+n] = torch.sin(2*np.pi*2*xh + 1.0)
    else:
        v[:, G:G+n] = torch.randn(B, n, generator=gen, device="cuda", dtype=torch.float64)
        w[:, G:G+n] = torch.randn(B, n, generator=gen, device="cuda", dtype=torch.float64)
    adj = AdjointEnsemble(s, nsteps=nsteps, dt=dt, segment=3)
    adj.forward(u0)
    jtw = adj.backward(w).clone()
    rhs = (jtw[:, G:G+n] * v[:, G:G+n]).sum(dim=1)
    out = []
    for h in (1e-4, 1e-6, 1e-8):
        s.solve_fixed_dt(u0 + h * v, dt, nsteps); up = s.u.clone()
        s.solve_fixed_dt(u0 - h * v, dt, nsteps)
        lhs = (w[:, G:G+n] * (up[:, G:G+n] - s.u[:, G:G+n]) / (2*h)).sum(dim=1)
        rel = (lhs - rhs).abs() / torch.maximum(lhs.abs(), rhs.abs()).clamp_min(1e-30)
        out.append(float(rel.max()))
    print(f"B={B} n={n} steps={nsteps} variant={variant} smooth_v={smooth_v}: rel err at h=1e-4,1e-6,1e-8: " + " ".join(f"{e:.2e}" for e in out))

for n in (256, 1024, 8192):
    for variant in (0, 1):
        for sm in (True, False):
            check(8, n, 6, variant, sm)
