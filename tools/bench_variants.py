"""Device-resident throughput of the schemes around the headline configuration (65536 x 4096 ensemble, fixed dt,
20 steps after 3 warm-up steps): which of them run the whole SSPRK33 step in one launch, which take three (or six)
stage launches.  One JSON line per scheme (-> profiles/)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from pyshocks_b200.ensemble import EnsembleSolver  # noqa: E402

B, N, G = 65536, 4096, 3
h = 3.0 / N
x = (torch.arange(N + 2 * G, device="cuda", dtype=torch.float64) - G + 0.5) / N
rng = np.random.default_rng(0)
coef = torch.from_numpy(rng.uniform(0.2, 1.0, size=(B, 1))).cuda()
u0 = 0.3 + coef * torch.sin(2 * np.pi * x)[None, :]
vi = 1.0 + 0.3 * np.sin(2 * np.pi * (np.arange(N) + 0.5) / N + 0.3)
vel = np.concatenate([vi[N - G :], vi, vi[:G]])  # ghost cells = periodic images (what psk_ssprk33_step asks for)
CASES = [
    ("burgers", "rusanov", "periodic"), ("burgers", "rusanov", "dirichlet"), ("burgers", "godunov", "periodic"),
    ("burgers", "eo", "dirichlet"), ("burgers", "lf", "periodic"), ("burgers", "rusanov", "neumann"),
    ("advection", "godunov", "dirichlet"), ("continuity", "godunov", "dirichlet"), ("advection", "godunov", "periodic"),
    ("continuity", "godunov", "periodic"),
    # alpha = 0.995: nu = df ** (alpha - 1) per face (the scheme of the reference's burgers-adjoint driver: lf, Dirichlet)
    ("burgers", "lf", "dirichlet", 0.995), ("burgers", "rusanov", "dirichlet", 0.995), ("burgers", "lf", "dirichlet"),
]
if os.environ.get("PSK_LF_WPC"):  # windows per CTA of the Lax-Friedrichs cluster kernel (tuning switch)
    from pyshocks_b200 import _lib
    assert _lib.lib().psk_set_stage_variant(8000 + int(os.environ["PSK_LF_WPC"])) == 0
if os.environ.get("PSK_STEP_VARIANT"):  # CTA shape of the whole-step kernel (tuning switch)
    from pyshocks_b200 import _lib
    assert _lib.lib().psk_set_stage_variant(7000 + int(os.environ["PSK_STEP_VARIANT"])) == 0
ONLY = sys.argv[1:]  # e.g. "neumann": the cases whose equation / flux / boundary kind is named
for eq, flux, bc, *rest in CASES:
    alpha = rest[0] if rest else 1.0
    if ONLY and not any(o in (eq, flux, bc, f"alpha={alpha}") for o in ONLY):
        continue
    kw = {"velocity": vel} if eq != "burgers" else {}
    if alpha != 1.0:
        kw["nu"] = np.diff(-1.37 + 3.1 * (np.arange(N + 2 * G) - G + 0.5) / N) ** (alpha - 1.0)
    s = EnsembleSolver(equation=eq, flux=flux, rec="wenojs53", bc=bc, n=N, g=G, dx=h, eps=1e-12, batch=B, **kw)
    if bc in ("dirichlet", "neumann"):
        s.hp.set_ghost(np.zeros(2 * G) if bc == "neumann" else np.full(2 * G, 0.3))
    dt = torch.full((1,), 0.4 * h / 1.5, dtype=torch.float64, device="cuda")
    s.load(u0)
    s.solve_fixed_dt(None, dt, 3)
    torch.cuda.synchronize()
    l0 = s.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.solve_fixed_dt(None, dt, 20)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"equation": eq, "flux": flux, "bc": bc, "alpha": alpha, "launches_per_step": (s.launches - l0) / 20,
                      "ms_per_step": ms / 20, "cell_updates_per_s": B * N * 20 / (ms * 1e-3),
                      "finite": bool(torch.isfinite(s.u[:, G : G + N]).all())}), flush=True)
    del s
    torch.cuda.empty_cache()
