"""Time one ``advance`` of ForwardEuler / SSPRK33 / RK44 / CKRK45 through the package API on a 16384 x 4096 ensemble
(Burgers, Rusanov, WENO-JS 5, periodic; FAST math): the stage combines fused into the right-hand side kernel
(psk_rhs_axpby) against the reference's array expressions around a fused apply_operator launch.
One JSON line per stepper (-> profiles/)."""
import json
import sys
from functools import partial

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import pyshocks_b200 as ps  # noqa: E402
import pyshocks_b200.timestepping as ts  # noqa: E402
from pyshocks_b200 import burgers  # noqa: E402
from pyshocks_b200.reconstruction import make_reconstruction_from_name  # noqa: E402
from pyshocks_b200.scalar import PeriodicBoundary  # noqa: E402

import os  # noqa: E402

if os.environ.get("PSK_FAST_WPC"):  # warps per CTA of the specialised stage kernels (tuning switch)
    from pyshocks_b200 import _lib
    assert _lib.lib().psk_set_stage_variant(4000 + int(os.environ["PSK_FAST_WPC"])) == 0
B, N, G = 16384, 4096, 3
grid = ps.make_uniform_cell_grid(a=-1.5, b=1.5, n=N, nghosts=G)
scheme = burgers.Rusanov(rec=make_reconstruction_from_name("wenojs53"), alpha=1.0)
bc = PeriodicBoundary()
rng = np.random.default_rng(0)
coef = torch.from_numpy(rng.uniform(0.2, 1.0, size=(B, 1))).cuda()
u = 0.3 + coef * torch.sin(2 * np.pi * grid.x.cuda() / 3.0)[None, :]
dt = 0.4 * grid.h / 1.5


def run(stepper, reps: int = 5) -> float:
    """a step() loop: every advance starts from the result of the last one"""
    out = u
    for _ in range(3):
        out = ts.advance(stepper, dt, 0.0, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = ts.advance(stepper, dt, 0.0, out)
    e1.record()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out[:, G:-G]).all())
    return e0.elapsed_time(e1) / reps, out


fused_binding = ts._fused_binding
for name in ("ForwardEuler", "SSPRK33", "RK44", "CKRK45"):
    mk = lambda: getattr(ts, name)(predict_timestep=lambda t_, u_: dt,  # noqa: E731
                                   source=partial(ps.apply_operator, scheme, grid, bc), checkpoint=None)
    ms_f, out_f = run(mk())
    ts._fused_binding = lambda *a: None
    try:
        ms_u, out_u = run(mk())
    finally:
        ts._fused_binding = fused_binding
    diff = float((out_f[:, G:-G] - out_u[:, G:-G]).abs().max())
    print(json.dumps({"stepper": name, "rows": B, "cells": N, "ms_fused_combines": ms_f, "ms_array_combines": ms_u,
                      "cell_updates_per_s_fused": B * N / (ms_f * 1e-3), "speedup": ms_u / ms_f,
                      "max_abs_diff": diff}), flush=True)
