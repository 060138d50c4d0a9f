"""Scratch: time the adjoint stage variants and the reverse sweep (config-5 shape)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200 import _lib
from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver
batch, n = 4096, 8192
h = 3.0 / n
s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=3, dx=h, eps=1e-12, batch=batch)
x = torch.linspace(0, 1, s.nx, device="cuda", dtype=torch.float64)
u0 = 0.5 + torch.sin(2 * np.pi * x)[None, :] * torch.rand(batch, 1, device="cuda", dtype=torch.float64)
s.load(u0)
dt = torch.full((1,), 0.4 * h / 1.5, dtype=torch.float64, device="cuda")
hp = s.hp
p, lam, out = s.new_states(3)
p.copy_(torch.randn_like(p)); lam.copy_(torch.randn_like(p))
ref = None
for variant in (0, 6, 7, 5, 0):
    _lib.lib().psk_set_adjoint_variant(variant)
    for _ in range(3):
        hp.stage_adjoint(s.u, p, dt, 1.0, out, acc=p, c_acc=1.0 / 3.0, acc2=lam, c_acc2=0.75)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        hp.stage_adjoint(s.u, p, dt, 1.0, out, acc=p, c_acc=1.0 / 3.0, acc2=lam, c_acc2=0.75)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None:
        ref = out.clone()
    err = float((out - ref).abs().max() / ref.abs().max())
    print(json.dumps({"variant": variant, "ms_per_adjoint_stage": ms, "cell_stages_per_s": batch * n / (ms * 1e-3), "rel_diff_vs_variant2": err}))
_lib.lib().psk_set_adjoint_variant(0)
