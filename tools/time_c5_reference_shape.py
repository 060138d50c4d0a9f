"""BASELINE configs[4] (4096 x 8192) in the scheme of the reference's own burgers-adjoint driver
(drivers/burgers-adjoint.py:68-97, 408: global Lax-Friedrichs with alpha = 0.995 on Dirichlet rows) next to the bench's
shape (Rusanov, alpha = 1, periodic rows): forward + reverse sweep over `nsteps` steps, gradients/s scaled to the 1000
steps of the named config.  One JSON line per shape (-> profiles/)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver  # noqa: E402

B, N, G = 4096, 8192, 3
NSTEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 100
h = 3.0 / N
x = (torch.arange(N + 2 * G, device="cuda", dtype=torch.float64) - G + 0.5) / N
coef = torch.from_numpy(np.random.default_rng(0).uniform(0.2, 1.0, size=(B, 1))).cuda()
u0 = 0.3 + coef * torch.sin(2 * np.pi * x)[None, :]
xc = -1.5 + 3.0 * (np.arange(N + 2 * G) - G + 0.5) / N
SHAPES = [
    ("bench shape: rusanov, alpha = 1, periodic", dict(flux="rusanov", bc="periodic")),
    ("rusanov, alpha = 1, dirichlet", dict(flux="rusanov", bc="dirichlet")),
    ("reference driver's shape: lf, alpha = 0.995, dirichlet", dict(flux="lf", bc="dirichlet", nu=np.diff(xc) ** (0.995 - 1.0))),
]
for name, kw in SHAPES:
    s = EnsembleSolver(equation="burgers", rec="wenojs53", n=N, g=G, dx=h, eps=1e-12, batch=B, **kw)
    if kw["bc"] == "dirichlet":
        s.hp.set_ghost(np.full(2 * G, 0.3))
    adj = AdjointEnsemble(s, nsteps=NSTEPS, dt=0.4 * h / 1.6 / 1.05)
    adj.gradient_half_l2(u0)  # warm-up
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    adj.forward(u0)
    e[1].record()
    torch.cuda.synchronize()
    fwd_ms = e[0].elapsed_time(e[1])
    e[0].record()
    J, grad = adj.gradient_half_l2(u0)
    e[1].record()
    torch.cuda.synchronize()
    tot_ms = e[0].elapsed_time(e[1])
    print(json.dumps({"shape": name, "rows": B, "cells": N, "steps": NSTEPS, "segment": adj.segment,
                      "reverse_mode": adj.reverse_mode, "forward_ms": fwd_ms, "forward_plus_reverse_ms": tot_ms,
                      "gradients_per_s_at_1000_steps": B / (tot_ms * 1e-3 * 1000.0 / NSTEPS),
                      "adjoint_cell_updates_per_s": B * N * NSTEPS / ((tot_ms - fwd_ms) * 1e-3),
                      "finite": bool(torch.isfinite(grad).all())}), flush=True)
    del adj, s
    torch.cuda.empty_cache()
