"""Small run of the fused reverse step on Dirichlet rows (psk_ssprk33_step_adjoint_bc, reverse_step_kernel<C, .., DIR>)
for compute-sanitizer (memcheck / racecheck): every run length, rows shorter than one window, rows whose last window
is short, per-row boundary data that differ at the three stage times, and a whole AdjointEnsemble sweep."""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from pyshocks_b200 import _lib as L  # noqa: E402
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver  # noqa: E402

G = 3
for n, variant in ((16, 0), (74, 12), (300, 16), (1000, 20), (1300, 24), (4300, 0), (8192, 0)):
    B = 3
    x = np.linspace(0, 6.28, n + 2 * G)
    u0 = torch.from_numpy(0.5 + np.sin(x)[None, :].repeat(B, 0) * np.array([[1.0], [0.5], [0.2]])).cuda()
    s = EnsembleSolver(batch=B, equation="burgers", flux="rusanov", rec="wenojs53", bc="dirichlet", n=n, g=G, dx=3.0 / n, eps=1e-12)
    assert L.lib().psk_set_reverse_variant(variant) == 0
    u, p, out, k1, k2 = s.new_states(5)
    u.copy_(u0)
    p[:, G : G + n] = 1.0
    g3 = torch.from_numpy(np.random.default_rng(n).uniform(-0.5, 0.5, size=(3, B, 2 * G))).cuda()
    dt = torch.full((1,), 0.3 * (3.0 / n) / 1.5, dtype=torch.float64, device="cuda")
    assert s.hp.reverse_step_fused(u, p, dt, out, stages=(k1, k2), ghosts=g3)
    assert bool(torch.isfinite(out[:, G : G + n]).all())
    L.lib().psk_set_reverse_variant(0)
    s.hp.set_ghost(np.full((B, 2 * G), 0.3))
    adj = AdjointEnsemble(s, nsteps=5, dt=float(dt), segment=2)
    assert adj.fused_reverse
    J, grad = adj.gradient_half_l2(u0)
    assert bool(torch.isfinite(grad).all())
torch.cuda.synchronize()
print("sanitize r2c target done")
