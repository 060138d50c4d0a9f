"""A/B of the specialised stage kernel's layout variants on the GPU box (scratch tool, not product).

For every variant (v < 1000: psk_set_stage_variant(5000 + layout) with three stage launches per step;
v >= 1000: the whole-step kernel psk_set_stage_variant(7000 + v - 1000)):
  * parity against the C oracle on a small ensemble (the checker), difference from layout 0,
    bitwise shift equivariance, the fused max-|u| path (adaptive solve) and the other Burgers
    fluxes of the specialised kernel;
  * throughput of the BASELINE configs[2] ensemble (65536 x 4096, fixed dt), best of 3 x 20 steps.
Writes one JSON line per variant to gpurun_out/ab_stage.jsonl and prints the ranking.
"""
from __future__ import annotations

import json
import os
import pathlib
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle.c_oracle import COracle  # noqa: E402  (checker)
from pyshocks_b200 import _lib  # noqa: E402
from pyshocks_b200.ensemble import EnsembleSolver  # noqa: E402

G = 3
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def ic(batch: int, n: int, seed: int) -> torch.Tensor:
    gen = torch.Generator(device="cuda").manual_seed(seed)
    c = torch.rand(batch, 1, generator=gen, device="cuda", dtype=torch.float64) - 0.5
    xh = ((torch.arange(n + 2 * G, device="cuda", dtype=torch.float64) - G + 0.5) / n)[None, :]
    u = c.repeat(1, n + 2 * G)
    for k in range(1, 5):
        a = torch.rand(batch, 1, generator=gen, device="cuda", dtype=torch.float64) / k
        ph = 2 * np.pi * torch.rand(batch, 1, generator=gen, device="cuda", dtype=torch.float64)
        u += a * torch.sin(2 * np.pi * k * xh + ph)
    return u


def solver(batch: int, n: int, **kw) -> EnsembleSolver:
    args = dict(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=G, dx=3.0 / n,
                eps=1e-12, batch=batch, math="fast")
    args.update(kw)
    return EnsembleSolver(**args)


def set_variant(v: int) -> None:
    """v < 1000: stage-kernel code (5000 + v), whole-step kernel off; v >= 1000: whole-step kernel
    7000 + (v - 1000) on top of the default stage kernel"""
    lib = _lib.lib()
    if v >= 1000:
        assert lib.psk_set_stage_variant(5002) == 0
        assert lib.psk_set_stage_variant(7000 + v - 1000) == 0
    else:
        assert lib.psk_set_stage_variant(7000) == 0
        rc = lib.psk_set_stage_variant(5000 + v)
        assert rc == 0, (v, rc)


def parity(v: int, base: dict) -> dict:
    res = {}
    # (1) small ensemble against the C oracle and against layout 0
    B, n, nsteps = 8, 4096, 15
    u0 = ic(B, n, seed=3)
    dt = 0.4 * (3.0 / n) / float(u0.abs().max())
    s = solver(B, n)
    s.solve_fixed_dt(u0, dt, nsteps)
    a = s.u[:, G : G + n].clone()
    if "oracle" not in base:
        co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=G, batch=B,
                     dx=3.0 / n, eps=1e-12)
        base["oracle"] = torch.from_numpy(co.solve_fixed_dt(u0.cpu().numpy(), dt, nsteps)[:, G : G + n]).cuda()
    ref = base["oracle"]
    res["err_oracle"] = float((a - ref).abs().max() / ref.abs().max())
    if "layout0" not in base:
        base["layout0"] = a
    res["diff_layout0"] = float((a - base["layout0"]).abs().max() / ref.abs().max())
    # (2) bitwise shift equivariance (k = 1237 as in tests/test_gpu_properties.py)
    u0s = u0.clone()
    u0s[:, G : G + n] = torch.roll(u0[:, G : G + n], 1237, dims=1)
    s2 = solver(B, n)
    s2.solve_fixed_dt(u0s, dt, nsteps)
    res["shift_bitwise"] = bool(torch.equal(torch.roll(a, 1237, dims=1), s2.u[:, G : G + n]))
    # (3) odd row length + fused max path: adaptive solve against layout 0's
    B3, n3 = 16, 1000
    u3 = ic(B3, n3, seed=11)
    s3 = solver(B3, n3)
    r3 = s3.solve_adaptive(u3, theta=0.5, tfinal=0.05, cfl_scale=0.5 * (3.0 / n3))
    key = "adaptive0"
    if key not in base:
        base[key] = (s3.u.clone(), int(r3.steps))
    res["adaptive_steps"] = int(r3.steps)
    res["adaptive_diff_layout0"] = float((s3.u[:, G : G + n3] - base[key][0][:, G : G + n3]).abs().max())
    res["adaptive_same_steps"] = bool(int(r3.steps) == base[key][1])
    # (4) every other flux of the specialised kernel, one step each, against layout 0
    worst = 0.0
    for flux in ("lf", "godunov", "eo"):
        s4 = solver(4, 2048, flux=flux)
        u4 = ic(4, 2048, seed=5)
        s4.solve_fixed_dt(u4, 0.3 * (3.0 / 2048) / float(u4.abs().max()), 5)
        k4 = "flux_" + flux
        if k4 not in base:
            base[k4] = s4.u.clone()
        worst = max(worst, float((s4.u[:, G:-G] - base[k4][:, G:-G]).abs().max()))
    res["other_fluxes_diff_layout0"] = worst
    return res


def sm_clock() -> int:
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    except Exception:  # noqa: BLE001
        return -1


def main() -> None:
    """interleaved rounds over the candidates on ONE solver (same buffers), after a thermal warm-up"""
    t_start = time.time()
    codes = [0, 2, 1060, 1062, 1082]
    if os.environ.get("AB_CODES"):
        codes = [int(x) for x in os.environ["AB_CODES"].split(",")]
    batch, n = int(os.environ.get("AB_BATCH", "65536")), 4096
    rounds, steps = int(os.environ.get("AB_ROUNDS", "5")), int(os.environ.get("AB_STEPS", "30"))
    base: dict = {}
    par = {}
    for v in codes:
        set_variant(v)
        try:
            par[v] = parity(v, base)
        except Exception as exc:  # noqa: BLE001  (scratch tool: record and go on)
            par[v] = {"error": repr(exc)}
        print(v, json.dumps(par[v]), flush=True)
    s = solver(batch, n)
    x = torch.linspace(0, 1, s.nx, device="cuda", dtype=torch.float64)
    gen = torch.Generator(device="cuda").manual_seed(1)
    u0 = 0.5 + torch.sin(2 * np.pi * x)[None, :] * torch.rand(batch, 1, device="cuda", dtype=torch.float64, generator=gen)
    s.load(u0)
    del u0
    dt = 0.4 * (3.0 / n) / 1.5
    set_variant(codes[0])
    s.solve_fixed_dt(None, dt, 400)  # ~1.5 s of load: clocks and temperature settle
    torch.cuda.synchronize()
    samples: dict[int, list[float]] = {v: [] for v in codes}
    clocks: dict[int, list[int]] = {v: [] for v in codes}
    for r in range(rounds):
        order = codes if r % 2 == 0 else codes[::-1]
        for v in order:
            set_variant(v)
            s._fused = None  # (the solver caches whether psk_ssprk33_step covers its scheme)
            s.solve_fixed_dt(None, dt, 2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.solve_fixed_dt(None, dt, steps)
            e1.record()
            clocks[v].append(sm_clock())
            torch.cuda.synchronize()
            samples[v].append(batch * n * steps / (e0.elapsed_time(e1) * 1e-3))
    assert bool(torch.isfinite(s.u).all())
    set_variant(2)
    rows = []
    with open(OUT / "ab_stage.jsonl", "w") as fh:
        for v in codes:
            xs = sorted(samples[v])
            row = {"variant": v,
                   "median": xs[len(xs) // 2], "min": xs[0], "max": xs[-1],
                   "hbm_frac_64B_median": xs[len(xs) // 2] * 64 / 6545.9e9, "sm_mhz": clocks[v], **par[v]}
            rows.append(row)
            fh.write(json.dumps(row) + "\n")
    rows.sort(key=lambda r: -r["median"])
    for r in rows:
        print(f'{r["variant"]:4d}  median {r["median"]:.4g}  [{r["min"]:.4g}, {r["max"]:.4g}]  frac {r["hbm_frac_64B_median"]:.3f}  '
              f'clk {r["sm_mhz"]}  err {r.get("err_oracle")}  shift {r.get("shift_bitwise")}  {r.get("error", "")}')
    print("wall", round(time.time() - t_start, 1))


if __name__ == "__main__":
    main()
