"""Small run of the whole-step kernel and the shared-difference stage kernels for compute-sanitizer."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pyshocks_b200 import _lib
from pyshocks_b200.ensemble import EnsembleSolver
B, g = 3, 3
for n in (50, 300, 1000):
    u0 = torch.from_numpy(0.5 + np.sin(np.linspace(0, 6.28, n + 2 * g))[None, :].repeat(B, 0)).cuda()
    for flux in ("rusanov", "godunov", "eo"):
        for code in (7066, 7061, 7062, 7064, 7060, 7082, 7000):
            _lib.lib().psk_set_stage_variant(code)
            s = EnsembleSolver(equation="burgers", flux=flux, rec="wenojs53", bc="periodic", n=n, g=g, dx=3.0 / n, eps=1e-12, batch=B)
            s.solve_fixed_dt(u0, 1e-3, 3)
            s.solve_adaptive(u0, theta=0.9, tfinal=0.004, cfl_scale=0.5 * 3.0 / n, check_every=1)
            assert bool(torch.isfinite(s.u[:, g:-g]).all())
_lib.lib().psk_set_stage_variant(7066)
torch.cuda.synchronize()
print("sanitize_step done")
