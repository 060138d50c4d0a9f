#!/usr/bin/env python
"""Benchmark of the pyshocks hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[2], the configuration the metric "WENO5 Burgers
cell-updates/s" is quoted on): an ensemble of B = 65536 independent inviscid Burgers problems x
N = 4096 cells, fp64, WENO-JS5 + Rusanov (LLF) + SSPRK33, periodic, random smooth initial data,
one shared fixed dt at CFL 0.4.  One "step" = one full SSPRK33 step of every cell of the ensemble
(one launch of the whole-step kernel psk_ssprk33_step per rank).  With N GPUs the B rows are
block-partitioned, B / N per rank (STRONG scaling, no collective in the data path); the weak
figure (B rows per GPU) is reported beside it as `weak`.

The same JSON line carries the other halves of the metric as sub-records, each measured in the
same process right after the headline:

  `slab`     BASELINE configs[3]: ONE periodic Burgers grid of 2^30 cells, slab-decomposed over
             the N ranks, ghost cells exchanged through NVLink peer memory;
  `adjoint`  BASELINE configs[4]: B = 4096 x N = 8192 ensemble (rows sharded over the ranks),
             1000 fixed-dt steps forward with a device tape + the reverse sweep: gradients/s
             (`adjoint.dirichlet`: the same ensemble on Dirichlet rows, 100 steps; `adjoint.reference_driver_scheme`:
             the scheme of drivers/burgers-adjoint.py itself -- global Lax-Friedrichs, alpha = 0.995, Dirichlet rows);
  `parity`   64 sampled rows of the timed ensemble state against the C restatement of the
             reference (oracle/psk_oracle.c) on the identical initial rows.

`--impl reference` times the CPU restatement of the reference path (oracle/psk_oracle.c, all host
threads) on a bounded sample of the same workload; the reference itself is Python-on-JAX and JAX is
not installable here.  Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "WENO5 Burgers cell-updates/s"
UNIT = "cell-updates/s"
N_CELLS = 4096
GHOSTS = 3
BATCH = 65536
DOMAIN = (-1.5, 1.5)
EPS = 1.0e-12
CFL = 0.4
ALGO_BYTES_PER_CELL_UPDATE = 64.0  # 16 + 24 + 24 B over the three stages
ALGO_FLOPS_PER_CELL_UPDATE = 456.0  # 152 per cell-stage, divisions counted as 1
ADJ_ALGO_BYTES = 144.0  # reverse sweep per cell-step, per-stage streaming design
# rows of the ensemble the CPU arm advances per step (0.55 GB: not cache resident); PSK_BENCH_CPU_ROWS: tests only
CPU_SAMPLE_ROWS = int(os.environ.get("PSK_BENCH_CPU_ROWS", "16384"))
PARITY_ROWS = 64
PARITY_TOL = 1.0e-12
SMS, FP64_LANES_PER_SM = 148, 64


def measured_peaks() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ensemble_coefficients(batch: int, seed: int) -> np.ndarray:
    """c_b, A_{b,1..4}, phi_{b,1..4} of u0_b(x) = c_b + sum_k A_bk sin(2 pi k xhat + phi_bk)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-0.5, 0.5, size=(batch, 1))
    amp = rng.uniform(0.0, 1.0, size=(batch, 4)) / np.arange(1, 5)
    phi = rng.uniform(0.0, 2.0 * np.pi, size=(batch, 4))
    return np.concatenate([c, amp, phi], axis=1)


def host_initial_condition(coef: np.ndarray, n: int, g: int) -> np.ndarray:
    nx = n + 2 * g
    xhat = (np.arange(nx) - g + 0.5) / n
    u = np.repeat(coef[:, :1], nx, axis=1)
    for k in range(4):
        u += coef[:, 1 + k : 2 + k] * np.sin(2.0 * np.pi * (k + 1) * xhat[None, :] + coef[:, 5 + k : 6 + k])
    return u


def device_initial_condition(coef, n: int, g: int, nx: int, dev):
    import torch

    xhat = ((torch.arange(nx, device=dev, dtype=torch.float64) - g + 0.5) / n)[None, :]
    u0 = coef[:, :1].repeat(1, nx)
    for k in range(4):
        u0 += coef[:, 1 + k : 2 + k] * torch.sin(2.0 * np.pi * (k + 1) * xhat + coef[:, 5 + k : 6 + k])
    return u0


def shard(total: int, rank: int, world: int) -> tuple[int, int]:
    base, extra = divmod(total, world)
    return rank * base + min(rank, extra), base + (1 if rank < extra else 0)


# {{{ CPU arm: ONE protocol for `--impl reference` and for `cpu_baseline`


def host_threads() -> int:
    return len(os.sched_getaffinity(0))


def cpu_port_steps(steps: int, warmup: int, threads: int | None = None) -> dict:
    """`warmup` + `steps` SSPRK33 steps of a bounded sample (CPU_SAMPLE_ROWS rows of the ensemble, all
    N_CELLS cells) on the C restatement, OpenMP over rows with an EXPLICIT thread count
    (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers)."""
    from oracle import c_oracle
    from oracle.c_oracle import COracle

    want = host_threads() if threads is None else threads
    used = c_oracle.set_threads(want)
    h = (DOMAIN[1] - DOMAIN[0]) / N_CELLS
    rows = CPU_SAMPLE_ROWS
    u = host_initial_condition(ensemble_coefficients(rows, 20261017), N_CELLS, GHOSTS)
    dt = CFL * h / np.abs(u).max()
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS, g=GHOSTS,
                 batch=rows, dx=h, eps=EPS)
    for _ in range(warmup):
        u = co.solve_fixed_dt(u, dt, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        u = co.solve_fixed_dt(u, dt, 1)
    wall = time.perf_counter() - t0
    return {
        "value": rows * N_CELLS * steps / wall, "unit": UNIT, "cores": used, "kind": "port",
        "sample": f"{rows} of {BATCH} rows x {N_CELLS} cells, {steps} SSPRK33 steps after {warmup} warm-up, one step per "
                  f"call (C restatement oracle/psk_oracle.c, OpenMP over rows, {used} threads set explicitly, {wall:.1f} s)",
        "seconds": wall,
    }


def run_reference(args: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_port_steps(args.steps, args.warmup)
    value = res["value"]
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "reference_arm": "CPU restatement of the reference path (oracle/psk_oracle.c); the reference is Python on "
                         "JAX and JAX is not installable here",
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# }}}


def measure_fp64_peak(dev) -> dict:
    """DFMA throughput of this GPU (MEASURED_PEAKS.json has no fp64 figure): psk_dfma_probe, best of 6."""
    import torch

    from pyshocks_b200 import _lib as L

    ctas, iters = 148 * 16, 2048
    out = torch.empty(ctas * 256, dtype=torch.float64, device=dev)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check("psk_dfma_probe", L.lib().psk_dfma_probe(L.ptr(out), ctas, iters, L.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * 64 * iters * ctas * 256 / (e0.elapsed_time(e1) * 1e-3))
    return {"tflops": best / 1e12,
            "how": "psk_dfma_probe in this run: 16 independent DFMA chains per thread, 2368 CTAs x 256 threads, best of 6"}


def sass_counts() -> dict:
    """Static FP64 instruction counts of the hot kernels (tools/sass_counts.py, regenerated by every
    build); `fresh` says whether they were taken from the library this run loaded."""
    f = ROOT / "profiles" / "sass_counts.json"
    if not f.exists():
        return {"kernels": {}, "fresh": False}
    d = json.loads(f.read_text())
    try:
        import hashlib

        lib = ROOT / "pyshocks_b200" / "csrc" / "libpsk.so"
        d["fresh"] = hashlib.sha1(lib.read_bytes()).hexdigest() == d.get("library_sha1")
    except OSError:
        d["fresh"] = False
    return d


def workload_config(n_gpus: int) -> dict:
    """Identical for both arms (the driver compares it)."""
    return {
        "workload": f"batched Burgers ensemble B={BATCH} x N={N_CELLS} cells, fp64, WENO-JS5 + Rusanov(LLF) "
                    f"+ SSPRK33, periodic, fixed dt at CFL {CFL} (BASELINE.json configs[2])",
        "batch": BATCH, "cells": N_CELLS, "ghosts": GHOSTS,
        "sharding": f"ensemble rows block-partitioned over {n_gpus} rank(s), B / N rows each, no data-path collective",
        "l2": "state arrays (2.15 GB / N per rank) are larger than the 126 MB L2 at every N <= 8; no flush needed",
    }


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md), through
    NVML from a background thread (every 2 ms); falls back to one nvidia-smi query."""

    REASONS = {
        "hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
        "hw_power_brake_slowdown": 0x80,
    }

    def __init__(self, index: int) -> None:
        import threading

        self.index = index
        self.sm: list[float] = []
        self.mask = 0
        self.max_mhz = None
        self.power: list[float] = []
        self._stop = threading.Event()
        self._pause = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._nvml = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self) -> None:
        nv = self._nvml
        if nv is None:
            return
        while not self._stop.is_set():
            if not self._pause.is_set():
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.002)

    def pause(self, on: bool) -> None:
        (self._pause.set if on else self._pause.clear)()

    def stop(self) -> dict:
        self._stop.set()
        self._thread.join(timeout=2)
        if self._nvml is None or not self.sm:
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = (float(x) for x in out.strip().split(","))
                return {"sm_mhz": a, "sm_max_mhz": b, "samples": 1, "reasons": ["sampled after the run (no NVML)"]}
            except Exception:  # noqa: BLE001
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock query unavailable"]}
        reasons = sorted(name for name, bit in self.REASONS.items() if self.mask & bit)
        return {
            "sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz,
            "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm), "reasons": reasons,
        }


class Ctx:
    """One rank of the bench: device, process group, barrier, max-over-ranks reduction."""

    def __init__(self) -> None:
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.fp64: dict | None = None  # DFMA peak measured in this run (rank 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self) -> None:
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values: list[float]) -> list[float]:
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def gather(self, values: list[float]) -> list[list[float]]:
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return [[float(x) for x in t]]
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[float(x) for x in o] for o in out]

    def close(self) -> None:
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(ctx: Ctx, fn, steps: int) -> float:
    """max over ranks of the CUDA-event time (ms) of `fn(steps)`, bracketed by barrier + synchronize"""
    torch = ctx.torch
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(steps)
    e1.record()
    ctx.barrier()
    return ctx.max_over_ranks([e0.elapsed_time(e1)])[0]


# {{{ headline: the ensemble (configs[2])


def measure_ensemble(ctx: Ctx, args: argparse.Namespace) -> dict:
    torch = ctx.torch
    from pyshocks_b200.ensemble import EnsembleSolver

    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    first, rows = shard(args.batch, rank, world)
    h = (DOMAIN[1] - DOMAIN[0]) / N_CELLS
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS,
                            g=GHOSTS, dx=h, eps=EPS, batch=rows, math="fast", device=dev)
    nx = solver.nx
    # synthetic initial data of the named shape: coefficients drawn on the host for the WHOLE ensemble
    # (the same rows whatever N), this rank's block evaluated on the device
    coef = torch.from_numpy(ensemble_coefficients(args.batch, 20261017)[first : first + rows]).to(dev)
    u0 = device_initial_condition(coef, N_CELLS, GHOSTS, nx, dev)
    umax = u0.abs().max()
    if world > 1:
        ctx.dist.all_reduce(umax, op=ctx.dist.ReduceOp.MAX)
    dt_host = CFL * h / float(umax)
    dt = torch.full((1,), dt_host, dtype=torch.float64, device=dev)
    solver.load(u0)
    # the rows the parity check follows: their INITIAL state goes to the host now
    prow = torch.linspace(0, rows - 1, min(PARITY_ROWS, rows), device=dev).round().long()
    u0_sample = u0[prow].cpu().numpy()

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps
    solver.solve_fixed_dt(None, dt, args.warmup)
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    launches0 = solver.launches
    ms_total = timed_steps(ctx, lambda k: solver.solve_fixed_dt(None, dt, k), args.steps)
    launches = solver.launches - launches0
    if sampler is not None:
        sampler.pause(True)
    finite = bool(torch.isfinite(solver.u).all())

    # ---- parity: the sampled rows after W + K steps against the C restatement on the identical initial rows
    from oracle import c_oracle
    from oracle.c_oracle import COracle

    c_oracle.set_threads(max(1, host_threads() // world))
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS, g=GHOSTS,
                 batch=u0_sample.shape[0], dx=h, eps=EPS)
    t0 = time.perf_counter()
    ref = co.solve_fixed_dt(u0_sample, dt_host, args.warmup + args.steps)
    parity_s = time.perf_counter() - t0
    got = solver.u[prow].cpu().numpy()
    i = slice(GHOSTS, GHOSTS + N_CELLS)
    rel = float(np.abs(got[:, i] - ref[:, i]).max() / np.abs(ref[:, i]).max())
    parity_rel = ctx.max_over_ranks([rel if np.isfinite(rel) else 1e300])[0]

    # ---- host <-> device copy bandwidth with every rank copying at once (what bounds e2e at N > 1)
    host_in = torch.empty((rows, nx), dtype=torch.float64, pin_memory=True)
    host_out = torch.empty((rows, nx), dtype=torch.float64, pin_memory=True)
    host_in.copy_(u0)
    del u0
    bytes_state = rows * nx * 8
    copy_ms = []
    for src, dst in ((host_in, solver.k2), (solver.k2, host_out)):
        dst.copy_(src, non_blocking=True)  # warm-up
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        ctx.barrier()
        copy_ms.append(e0.elapsed_time(e1))
    per_rank_gbs = ctx.gather([bytes_state / (copy_ms[0] * 1e6), bytes_state / (copy_ms[1] * 1e6)])
    copy_bound_ms = ctx.max_over_ranks([max(copy_ms)])[0]

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    solver.solve_fixed_dt_host(host_in, host_out, dt, 1)  # warm-up of the copy streams
    ctx.barrier()
    if sampler is not None:
        sampler.pause(False)
    t0 = time.perf_counter()
    solver.solve_fixed_dt_host(host_in, host_out, dt, args.steps)
    ctx.barrier()
    e2e_s = ctx.max_over_ranks([time.perf_counter() - t0])[0]
    clocks = sampler.stop() if sampler is not None else None
    groups = max(1, min(64, rows // 1024))
    del host_in, host_out, solver
    torch.cuda.empty_cache()
    if rank != 0:
        return {}

    cells_per_step = args.batch * N_CELLS
    value = cells_per_step * args.steps / (ms_total * 1e-3)
    peak, peak_src = measured_peaks()
    per_gpu = value / world
    per_step = launches / max(args.steps, 1)  # 1: whole-step kernel, 3: one launch per stage
    whole = per_step < 2
    fp64 = measure_fp64_peak(dev)
    ctx.fp64 = fp64
    sc = sass_counts()
    k = sc["kernels"]
    if whole and k.get("step_fused"):
        per_update = k["step_fused"][0]["fp64"] * 32 / 172  # per lane / (172 emitted cells / 32 lanes)
        kernel = k["step_fused"][0]["name"]
    elif k.get("stage1"):
        per_update = sum(k[s][0]["fp64"] for s in ("stage1", "stage2", "stage3")) * 32 / 120
        kernel = "psk::stage_warp_fast_share_kernel x 3"
    else:  # no count file: the figures of round 1
        per_update, kernel = (922 * 32 / 172 if whole else 654 * 32 / 120), "(static counts missing)"
    # FP64 pipe: every FP64-pipe warp instruction occupies one issue slot of the 64 lanes/clk/SM pipe, whatever it
    # is (DFMA, DMUL, DADD); counted as one DFMA slot = 2 flop, against the DFMA rate measured in this run
    slot_tflops = per_gpu * per_update * 2.0 / 1e12
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        tj = json.loads(tf.read_text())
        per_cell = tj.get("step_kernel", {}).get("dram_bytes_per_cell_measured") if whole else None
        if per_cell is not None:
            traffic = per_cell * rows * N_CELLS  # per launch on this rank
    achieved_hbm = per_gpu * ALGO_BYTES_PER_CELL_UPDATE / 1e9
    agg = [sum(r[0] for r in per_rank_gbs), sum(r[1] for r in per_rank_gbs)]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "rows_per_gpu": rows, "finite": finite,
        "parity": {
            "max_rel": parity_rel, "tol": PARITY_TOL, "ok": bool(parity_rel <= PARITY_TOL),
            "what": f"{min(PARITY_ROWS, rows)} evenly spaced rows per rank of the timed device state after "
                    f"{args.warmup + args.steps} steps (FAST whole-step kernel) against oracle/psk_oracle.c advanced from "
                    f"the identical initial rows; max |diff| / max |ref| over the interior, max over ranks",
            "oracle_seconds": parity_s,
        },
        # the BINDING roofline of this path is the FP64 pipe (ncu: pipe 84 % busy, DRAM 16 %), so `roofline` is
        # reported against it; the 64-B HBM figure of SURVEY.md 8(d) follows as `roofline_hbm`
        "roofline": {
            "bound": "fp64", "achieved": slot_tflops, "peak": fp64["tflops"], "unit": "TFLOP/s",
            "frac": slot_tflops / fp64["tflops"], "traffic": traffic,
            "kernel": kernel,
            "definition": "executed FP64-pipe warp instructions per cell-update (static SASS count of the straight-line "
                          "kernel per lane / cells a lane emits: profiles/sass_counts.json, regenerated by every build) x "
                          "cell-updates/s x 2 flop per pipe slot, against the DFMA peak measured in this run",
            "executed_fp64_instr_per_cell_update": per_update, "sass_counts_fresh": bool(sc.get("fresh")),
            "peak_source": fp64["how"],
            "algorithmic_flops": {"per_cell_update": ALGO_FLOPS_PER_CELL_UPDATE,
                                  "tflops": per_gpu * ALGO_FLOPS_PER_CELL_UPDATE / 1e12,
                                  "note": "SURVEY.md 8(d) count of the reference's arithmetic (divisions = 1 flop); the "
                                          "FMA-fused, re-associated kernel executes fewer operations, so this can exceed the peak"},
            "pipe_frac_at_sampled_clock": (per_gpu * per_update / (SMS * FP64_LANES_PER_SM * clocks["sm_mhz"] * 1e6)
                                           if clocks and clocks.get("sm_mhz") else None),
            "avg_launch_ms": ms_total / max(launches, 1),
        },
        "roofline_hbm": {
            "bound": "hbm", "achieved": achieved_hbm, "peak": peak, "unit": "GB/s", "frac": achieved_hbm / peak,
            "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL_UPDATE / per_step * rows * N_CELLS,
            "note": "64 algorithmic bytes per cell-update (three streaming stages, SURVEY.md 8d); the whole-step kernel "
                    "moves 15.7 B of them (traffic: ncu dram bytes per cell of one launch, profiles/traffic.json, x the cells of "
                    "this launch)",
        },
        "e2e": {
            "value": cells_per_step * args.steps / e2e_s, "unit": UNIT,
            "h2d_bytes_per_step": bytes_state * world / args.steps, "d2h_bytes_per_step": bytes_state * world / args.steps,
            "call": f"EnsembleSolver.solve_fixed_dt_host(pinned host in, pinned host out, dt, K={args.steps}) per rank: "
                    f"{groups} row blocks, upload / K steps / download pipelined over 4 streams",
            "seconds": e2e_s,
            "amortisation": f"ONE upload and ONE download of the state per call, amortised over K = {args.steps} steps "
                            f"(at K = 1 the call is PCIe-bound: 2 x {bytes_state / 1e9:.2f} GB per rank per step)",
            "h2d_gbs_per_rank": [r[0] for r in per_rank_gbs], "d2h_gbs_per_rank": [r[1] for r in per_rank_gbs],
            "h2d_gbs": agg[0], "d2h_gbs": agg[1],
            "copy_bound_seconds": copy_bound_ms * 1e-3,
            "note": "h2d / d2h: one whole-state pinned copy per direction with every rank copying at the same time (CUDA "
                    "events, per rank); copy_bound_seconds = the slower direction alone on the slowest rank: a floor of "
                    "`seconds` that no kernel can lower",
        },
        "gpu_launches": launches,
        "clocks": clocks,
    }
    return line


def measure_weak(ctx: Ctx, args: argparse.Namespace) -> dict:
    """B rows PER GPU (round 1's headline): no collective, so it only shows that N ranks do not interfere."""
    torch = ctx.torch
    from pyshocks_b200.ensemble import EnsembleSolver

    h = (DOMAIN[1] - DOMAIN[0]) / N_CELLS
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS,
                            g=GHOSTS, dx=h, eps=EPS, batch=args.batch, math="fast", device=ctx.dev)
    coef = torch.from_numpy(ensemble_coefficients(args.batch, 20261017 + ctx.rank)).to(ctx.dev)
    u0 = device_initial_condition(coef, N_CELLS, GHOSTS, solver.nx, ctx.dev)
    dt = torch.full((1,), CFL * h / 3.0, dtype=torch.float64, device=ctx.dev)  # |u0| < 0.5 + sum 1/k < 3
    solver.load(u0)
    del u0
    solver.solve_fixed_dt(None, dt, args.warmup)
    ms = timed_steps(ctx, lambda k: solver.solve_fixed_dt(None, dt, k), args.steps)
    del solver
    torch.cuda.empty_cache()
    return {"value": args.batch * N_CELLS * ctx.world * args.steps / (ms * 1e-3), "unit": UNIT,
            "rows_per_gpu": args.batch, "ms_per_step": ms / args.steps, "scaling": "weak"}


# the other schemes of the path, each ONE launch per step since round 2: (equation, flux, boundary kind, alpha)
SCHEMES = [
    ("burgers", "godunov", "periodic", 1.0), ("burgers", "rusanov", "dirichlet", 1.0), ("burgers", "rusanov", "neumann", 1.0),
    ("burgers", "eo", "dirichlet", 1.0), ("burgers", "lf", "periodic", 1.0),
    ("burgers", "lf", "dirichlet", 0.995),  # the scheme of the reference's burgers-adjoint driver (drivers/burgers-adjoint.py:408)
    ("advection", "godunov", "periodic", 1.0), ("continuity", "godunov", "dirichlet", 1.0),
]


def measure_schemes(ctx: Ctx, args: argparse.Namespace) -> dict:
    """The configs[2] ensemble (same rows, same initial data, strong scaling) under the OTHER fluxes, boundary kinds and
    equations of the path: device-resident throughput, launches per step, and the first
    rows after W + K steps against the C restatement bound to the same scheme."""
    torch = ctx.torch
    from oracle import c_oracle
    from oracle.c_oracle import COracle
    from pyshocks_b200.ensemble import EnsembleSolver

    h = (DOMAIN[1] - DOMAIN[0]) / N_CELLS
    r0, rows = shard(args.batch, ctx.rank, ctx.world)
    nx = N_CELLS + 2 * GHOSTS
    coef_h = ensemble_coefficients(args.batch, 20261017)[r0 : r0 + rows]
    coef = torch.from_numpy(coef_h).to(ctx.dev)
    u0 = device_initial_condition(coef, N_CELLS, GHOSTS, nx, ctx.dev)
    nsample = min(8, rows)
    u0_sample = u0[:nsample].cpu().numpy()
    x = (np.arange(N_CELLS) + 0.5) / N_CELLS
    vi = 0.2 + np.sin(2.0 * np.pi * x + 0.3)  # changes sign; ghost cells = periodic images
    vel = np.concatenate([vi[N_CELLS - GHOSTS :], vi, vi[:GHOSTS]])
    ghost = np.array([0.31, -0.12, 0.05, -0.22, 0.17, 0.08])
    c_oracle.set_threads(max(1, host_threads() // ctx.world))
    out = []
    for eq, flux, bc, alpha in SCHEMES:
        kw: dict = {}
        if eq != "burgers":
            kw["velocity"] = vel
        if alpha != 1.0:  # nu = df ** (alpha - 1) per face (scalar.py:231-234)
            xc = DOMAIN[0] + (DOMAIN[1] - DOMAIN[0]) * (np.arange(nx) - GHOSTS + 0.5) / N_CELLS
            kw["nu"] = np.diff(xc) ** (alpha - 1.0)
        gh = None if bc == "periodic" else (ghost * (h if bc == "neumann" else 1.0))
        solver = EnsembleSolver(equation=eq, flux=flux, rec="wenojs53", bc=bc, n=N_CELLS, g=GHOSTS, dx=h, eps=EPS,
                                batch=rows, math="fast", device=ctx.dev, **kw)
        if gh is not None:
            solver.hp.set_ghost(gh)
        dt_host = CFL * h / 3.0 / (float(kw["nu"].max()) if "nu" in kw else 1.0)
        dt = torch.full((1,), dt_host, dtype=torch.float64, device=ctx.dev)
        solver.load(u0)
        solver.solve_fixed_dt(None, dt, args.warmup)
        l0 = solver.launches
        ms = timed_steps(ctx, lambda k, s=solver, d=dt: s.solve_fixed_dt(None, d, k), args.steps)
        launches = (solver.launches - l0) / args.steps
        co = COracle(equation=eq, flux=flux, rec="wenojs53", bc=bc, n=N_CELLS, g=GHOSTS, batch=nsample, dx=h, eps=EPS,
                     **kw)
        if gh is not None:
            co.set_ghost(gh)
        ref = co.solve_fixed_dt(u0_sample, dt_host, args.warmup + args.steps)
        got = solver.u[:nsample].cpu().numpy()
        i = slice(GHOSTS, GHOSTS + N_CELLS)
        rel = float(np.abs(got[:, i] - ref[:, i]).max() / np.abs(ref[:, i]).max())
        rel = ctx.max_over_ranks([rel if np.isfinite(rel) else 1e300])[0]
        out.append({"scheme": f"{eq}/{flux}/{bc}" + ("" if alpha == 1.0 else f"/alpha={alpha}"),
                    "value": args.batch * N_CELLS * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                    "launches_per_step": launches, "parity_max_rel": rel})
        del solver
        torch.cuda.empty_cache()
    del u0
    torch.cuda.empty_cache()
    return {"unit": UNIT, "rows_total": args.batch, "cells": N_CELLS, "steps": args.steps, "scaling": "strong",
            "parity": f"first {nsample} rows of every rank after W + K steps against oracle/psk_oracle.c bound to the same "
                      "scheme, max over ranks", "list": out}


# }}}

# {{{ single huge grid (configs[3])


def slab_window_check(slab, n_global: int, h: float, dt: float, nsteps: int, ctx: Ctx) -> dict:
    """Size-independent parity of the decomposed solve: the domain of dependence of a cell is 9 cells per
    step, so a window of the global grid advanced on its own by the C restatement (boundary kind NONE,
    analytic initial data) must reproduce the slab's cells at the window centre.  Checked at this rank's
    LAST cells -- the window reaches into the right neighbour's slab (or wraps around at N = 1), i.e.
    the values that travelled through the ghost-cell exchange."""
    from oracle import c_oracle
    from oracle.c_oracle import COracle

    halo = 9 * nsteps + 16
    core = 2048
    n_w = core + 2 * halo
    start = slab.first + slab.n_local - core // 2 - halo  # global index of the first window cell
    idx = (np.arange(start - GHOSTS, start + n_w + GHOSTS) % n_global).astype(np.float64)
    u = (0.5 + np.sin(2.0 * np.pi * (idx + 0.5) / n_global))[None, :]
    c_oracle.set_threads(1)
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="none", n=n_w, g=GHOSTS, batch=1, dx=h, eps=EPS)
    ref = co.solve_fixed_dt(u, dt, nsteps)[0, GHOSTS + halo : GHOSTS + halo + core // 2]
    got = slab.interior()[slab.n_local - core // 2 :].cpu().numpy()
    rel = float(np.abs(got - ref).max() / np.abs(ref).max())
    rel = ctx.max_over_ranks([rel if np.isfinite(rel) else 1e300])[0]
    return {"max_rel": rel, "tol": 1.0e-11, "ok": bool(rel <= 1.0e-11),
            "what": f"the last {core // 2} cells of every rank's slab after {nsteps} steps against oracle/psk_oracle.c on a "
                    f"{n_w}-cell window of the global grid around the slab edge (domain of dependence 9 cells per step); "
                    "initial data evaluated on the host (device sin differs by an ulp); max over ranks"}


def measure_slab(ctx: Ctx, args: argparse.Namespace, *, n_global: int, transport: str, steps: int, warmup: int,
                 graph: bool = False, check: bool = True) -> dict:
    torch = ctx.torch
    from pyshocks_b200.distributed import DistRing, PeerRing, PeerSlabSolver, SlabSolver

    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    h = (DOMAIN[1] - DOMAIN[0]) / n_global
    whole = transport in ("p2p-step", "p2p-step-fused")
    if transport == "nccl":
        slab = SlabSolver(n_global=n_global, ring=DistRing(), dx=h, device=dev)
    else:
        kw = {"fused_step": transport == "p2p-step-fused"} if whole else {}
        slab = PeerSlabSolver(n_global=n_global, rank=rank, world=world, dx=h, device=dev, whole_step=whole,
                              overlap=(transport == "p2p-overlap"), fused=(transport == "p2p"), **kw)
        if world > 1:
            slab.connect()
        else:
            slab.attach(PeerRing.local([slab.mem], 0))
    i = torch.arange(slab.first, slab.first + slab.n_local, device=dev, dtype=torch.float64)
    slab.load_interior(0.5 + torch.sin(2.0 * np.pi * (i + 0.5) / n_global))
    del i
    dt_host = CFL * h / 1.5
    dt = torch.full((1,), dt_host, dtype=torch.float64, device=dev)
    kw = {"graph": True} if (transport == "p2p" and graph) else {}
    warm = max(warmup, 6 if kw else 0)
    slab.solve_fixed_dt(dt, warm, **kw)
    l1 = getattr(slab, "launches", 0)
    ms = timed_steps(ctx, lambda k: slab.solve_fixed_dt(dt, k, **kw), steps)
    launches = (getattr(slab, "launches", 0) - l1) if transport != "nccl" else 3 * steps
    if transport != "nccl":
        slab.check()  # a ghost-cell wait that timed out would have produced garbage silently
    finite = bool(ctx.max_over_ranks([0.0 if bool(torch.isfinite(slab.interior()).all().item()) else 1.0])[0] == 0.0)
    parity = slab_window_check(slab, n_global, h, dt_host, warm + steps, ctx) if check else None
    n_local = slab.n_local
    if transport != "nccl":
        if world > 1:
            slab.close()
        else:
            slab.ring = None
            slab.solver = None
            slab._graph = None
            slab.mem.close()
    del slab
    torch.cuda.empty_cache()
    value = n_global * steps / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    achieved = value / world * ALGO_BYTES_PER_CELL_UPDATE / 1e9
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms / steps, "scaling": "strong", "cells": n_global, "cells_per_gpu": n_local,
        "workload": f"single periodic Burgers grid N={n_global} cells, slab-decomposed over {world} rank(s) "
                    "(BASELINE.json configs[3])",
        "halo_exchanges_per_step": 1 if whole else 3, "halo_cells_per_side": 9 if whole else 3,
        "transport": transport,
        "transport_detail": {
            "p2p": "exchange fused into the stage kernel: edge warps spin on local epoch flags, edge lanes store into the "
                   "neighbours' ghost slots over NVLink (1 launch per stage)",
            "p2p-overlap": "NVLink peer stores + epoch flags, slab edges on a high-priority stream overlapped with the interior",
            "p2p-serial": "NVLink peer stores + epoch flags, no overlap",
            "p2p-step": "whole SSPRK33 step in one launch on a slab with 9 ghost cells; one exchange per step "
                        "(psk_halo_wait -> psk_ssprk33_step -> psk_halo_push: NVLink peer stores + epoch flags)",
            "p2p-step-fused": "whole SSPRK33 step AND the 9-cell exchange in ONE launch: edge warps spin on local epoch "
                              "flags, the lanes that store the slab's outermost 9 cells also store them into the "
                              "neighbours' ghost slots over NVLink and raise their flags",
            "nccl": "NCCL send/recv pairs per stage"}[transport],
        "cuda_graph": bool(kw), "finite": finite, "parity": parity,
        "roofline_hbm": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src},
        "per_gpu_value": value / world,
        "gpu_launches": launches,
    }


# }}}

# {{{ adjoint (configs[4])


def adjoint_twin_check(dev, n: int, h: float, dt: float, u0_rows: np.ndarray, nsteps: int, kw: dict,
                       ghost: np.ndarray | None = None, flux: str = "rusanov", alpha: float = 1.0) -> dict:
    """The product's gradient of J = 1/2 ||u(T)||^2 on a few rows of the benchmarked ensemble over a few
    steps against reverse-mode differentiation of the reference arithmetic (oracle/torch_twin.py).
    `ghost`: (rows, 2 g) Dirichlet data (time-independent) -- Dirichlet rows instead of periodic ones."""
    import torch

    from oracle import pyshocks_oracle as po
    from oracle import torch_twin as tt
    from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

    rows = u0_rows.shape[0]
    grid = po.make_grid(DOMAIN[0], DOMAIN[1], n, GHOSTS)
    nu = None if alpha == 1.0 else np.diff(grid.x) ** (alpha - 1.0)  # grid.df ** (alpha - 1), scalar.py:231-232
    solver = EnsembleSolver(equation="burgers", flux=flux, rec="wenojs53", bc="periodic" if ghost is None else "dirichlet",
                            n=n, g=GHOSTS, dx=h, eps=EPS, batch=rows, device=dev, nu=nu)
    if ghost is not None:
        solver.hp.set_ghost(ghost)
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, **(kw | {"segment": 2}))
    _, grad = adj.gradient_half_l2(torch.from_numpy(u0_rows).to(dev))
    grad = grad.cpu().numpy()
    scheme = po.Scheme("burgers", flux, po.make_reconstruction("wenojs53"), alpha=alpha)
    xg = np.concatenate([grid.x[:GHOSTS], grid.x[-GHOSTS:]])
    worst = 0.0
    for b in range(rows):
        bc = po.Periodic() if ghost is None else po.Dirichlet(ga=lambda t, x, b=b: np.interp(x, xg, ghost[b]))
        u = torch.from_numpy(u0_rows[b]).clone().requires_grad_(True)
        x = u
        for _ in range(nsteps):
            x = tt.ssprk33_advance(lambda t_, y: tt.apply_operator(scheme, grid, bc, t_, y), dt, 0.0, x)
        (gb,) = torch.autograd.grad(0.5 * (x[grid.interior] ** 2).sum(), u)
        gb = gb.numpy()
        i = grid.interior
        worst = max(worst, float(np.abs(grad[b][i] - gb[i]).max() / np.abs(gb[i]).max()))
    # periodic rows of smooth data: 1e-12; Dirichlet rows whose boundary data (the row's mean) jump against the interior:
    # the suite's tolerance for non-smooth data (tests/test_gpu_adjoint.py tol_for: the derivative of the weights
    # amplifies round-off where beta ~ eps next to a jump, here in the cells at the row ends)
    tol = 1.0e-12 if ghost is None else 1.0e-9
    return {"max_rel": worst, "tol": tol, "ok": bool(worst <= tol),
            "what": f"dJ/du0 of {rows} rows of this ensemble over {nsteps} steps (same kernels, two-level tape) against "
                    "torch autograd through oracle/torch_twin.py (the stand-in for jax.jacfwd of the reference's advance)"
                    + ("" if ghost is None else "; Dirichlet rows: non-smooth at the row ends, tolerance of the suite's non-smooth cases")}


def measure_adjoint_dirichlet(ctx: Ctx, *, batch_total: int, n: int, nsteps: int, kw: dict, check: bool,
                              flux: str = "rusanov", alpha: float = 1.0) -> dict:
    """configs[4] on DIRICHLET rows -- the boundary kind of the reference's own burgers-adjoint template
    (drivers/burgers-adjoint.py:68-97) -- over `nsteps` steps with every state kept: whole-step forward launches
    (psk_ssprk33_step_bc) and one launch per reverse step (psk_ssprk33_step_adjoint_bc).  Same rows and initial data
    as the periodic record; every row's boundary data are its own mean value c_b (time-independent)."""
    torch = ctx.torch
    from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    first, rows = shard(batch_total, rank, world)
    h = (DOMAIN[1] - DOMAIN[0]) / n
    xc = DOMAIN[0] + h * (np.arange(n + 2 * GHOSTS) - GHOSTS + 0.5)
    nu = None if alpha == 1.0 else np.diff(xc) ** (alpha - 1.0)  # grid.df ** (alpha - 1), scalar.py:231-232
    solver = EnsembleSolver(equation="burgers", flux=flux, rec="wenojs53", bc="dirichlet", n=n, g=GHOSTS,
                            dx=h, eps=EPS, batch=rows, device=dev, nu=nu)
    coef_host = ensemble_coefficients(batch_total, 20261018)
    ghost_host = np.repeat(coef_host[first : first + rows, :1], 2 * GHOSTS, axis=1)
    solver.hp.set_ghost(ghost_host)
    coef = torch.from_numpy(coef_host[first : first + rows]).to(dev)
    u0 = device_initial_condition(coef, n, GHOSTS, solver.nx, dev)
    umax = u0.abs().max()
    if world > 1:
        ctx.dist.all_reduce(umax, op=ctx.dist.ReduceOp.MAX)
    dt = CFL * h / float(umax)
    if flux != "rusanov":
        kw = {k: v for k, v in kw.items() if k != "fused_reverse"}
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=1, **kw)
    small = AdjointEnsemble(solver, nsteps=2, dt=dt, segment=1, **kw)
    small.gradient_half_l2(u0)  # warm-up
    del small
    ctx.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    uT = adj.forward(u0)
    ev[1].record()
    pT = adj.lam1
    pT.zero_()
    pT[:, GHOSTS : GHOSTS + n] = uT[:, GHOSTS : GHOSTS + n]
    grad = adj.backward(pT)
    ev[2].record()
    ctx.barrier()
    fwd_ms, bwd_ms = ctx.max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])])
    out = {
        "workload": f"the same ensemble on Dirichlet rows (boundary data = the row's mean value), {flux} flux, alpha = {alpha}, "
                    f"{nsteps} steps, every state kept",
        "steps": nsteps, "forward_ms": fwd_ms, "reverse_ms": bwd_ms, "reverse_mode": adj.reverse_mode,
        "gradients_per_s_at_these_steps": batch_total / ((fwd_ms + bwd_ms) * 1e-3),
        "adjoint_cell_updates_per_s": batch_total * n * nsteps / (bwd_ms * 1e-3),
        "forward_cell_updates_per_s": batch_total * n * nsteps / (fwd_ms * 1e-3),
        "grad_finite": bool(torch.isfinite(grad).all()), "gpu_launches": adj.launches,
    }
    u0_rows = u0[:2].cpu().numpy()
    del adj, grad, uT, pT, u0, solver
    torch.cuda.empty_cache()
    if check and rank == 0:
        out["parity"] = adjoint_twin_check(dev, n, h, dt, u0_rows, 6, kw, ghost=ghost_host[:2], flux=flux, alpha=alpha)
    torch.cuda.empty_cache()
    return out


def measure_adjoint(ctx: Ctx, args: argparse.Namespace, *, batch_total: int, n: int, nsteps: int, check: bool = True) -> dict:
    torch = ctx.torch
    from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    first, rows = shard(batch_total, rank, world)
    h = (DOMAIN[1] - DOMAIN[0]) / n
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=GHOSTS,
                            dx=h, eps=EPS, batch=rows, device=dev)
    coef_host = ensemble_coefficients(batch_total, 20261018)
    coef = torch.from_numpy(coef_host[first : first + rows]).to(dev)
    u0 = device_initial_condition(coef, n, GHOSTS, solver.nx, dev)
    umax = u0.abs().max()
    if world > 1:
        ctx.dist.all_reduce(umax, op=ctx.dist.ReduceOp.MAX)
    dt = CFL * h / float(umax)
    kw = {}
    if os.environ.get("PSK_FUSED_REVERSE") is not None:  # A/B runs
        kw["fused_reverse"] = os.environ["PSK_FUSED_REVERSE"] == "1"
    if os.environ.get("PSK_FUSED_RECOMPUTE") is not None:
        kw["fused_recompute"] = os.environ["PSK_FUSED_RECOMPUTE"] == "1"
    seg = {"segment": int(os.environ["PSK_ADJ_SEGMENT"])} if os.environ.get("PSK_ADJ_SEGMENT") else {}
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, **kw, **seg)
    small = AdjointEnsemble(solver, nsteps=4, dt=dt, segment=2, **kw)
    small.gradient_half_l2(u0)  # warm-up of every kernel
    del small
    ctx.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    uT = adj.forward(u0)
    ev[1].record()
    pT = adj.lam1
    pT.zero_()
    pT[:, GHOSTS : GHOSTS + n] = uT[:, GHOSTS : GHOSTS + n]
    grad = adj.backward(pT)
    ev[2].record()
    ctx.barrier()
    fwd_ms, bwd_ms = ctx.max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])])
    grad_finite = bool(torch.isfinite(grad).all())
    segment, tape_states, launches = adj.segment, len(adj.chk) + len(adj.ring), adj.launches
    mode = getattr(adj, "reverse_mode", "recompute k1, k2 + 3 adjoint stage launches per reverse step")
    u0_rows = u0[:2].cpu().numpy()
    del adj, grad, uT, pT, u0, solver
    torch.cuda.empty_cache()
    parity = adjoint_twin_check(dev, n, h, dt, u0_rows, 6, kw) if (check and rank == 0) else None
    torch.cuda.empty_cache()
    peak, peak_src = measured_peaks()
    cells = batch_total * n
    adj_rate = cells * nsteps / (bwd_ms * 1e-3)
    achieved = adj_rate / world * ADJ_ALGO_BYTES / 1e9
    # FP64 pipe (the binding unit): executed FP64 instructions per cell-step of the reverse sweep = the fused reverse
    # kernel's (ncu, profiles/traffic.json) + the whole-step recomputation of the states inside a tape segment
    roof = None
    tf = ROOT / "profiles" / "traffic.json"
    if rank == 0 and tf.exists() and "1 launch" in mode:
        tj = json.loads(tf.read_text())
        fp64 = ctx.fp64 if ctx.fp64 is not None else measure_fp64_peak(dev)
        rev = tj.get("reverse_kernel", {}).get("fp64_thread_instructions_per_cell_executed")
        stp = tj.get("step_kernel", {}).get("fp64_thread_instructions_per_cell_executed")
        if rev and stp:
            per_cell = rev + stp * (segment - 1) / segment
            slots = adj_rate / world * per_cell * 2.0 / 1e12
            roof = {"bound": "fp64", "achieved": slots, "peak": fp64["tflops"], "unit": "TFLOP/s", "frac": slots / fp64["tflops"],
                    "executed_fp64_instr_per_cell_step": per_cell, "peak_source": fp64["how"],
                    "traffic_bytes_per_cell_step": tj["reverse_kernel"]["dram_bytes_per_cell_measured"],
                    "definition": "executed FP64 thread instructions per cell-step (ncu counters of the reverse kernel "
                                  "at this row length + the whole-step kernel's for the states recomputed inside a tape "
                                  "segment, profiles/traffic.json) x adjoint cell-updates/s x 2 flop per pipe slot, against "
                                  f"the DFMA peak measured in this run; ncu reports the pipe {tj['reverse_kernel'].get('fp64_pipe_pct', 0):.1f} % busy "
                                  "for the kernel alone"}
    try:
        dirichlet = measure_adjoint_dirichlet(ctx, batch_total=batch_total, n=n, nsteps=min(nsteps, 100), kw=kw, check=check)
    except Exception as exc:  # noqa: BLE001  (must not take the named record with it)
        import traceback

        traceback.print_exc()
        dirichlet = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
    try:  # the scheme of the reference's own driver (drivers/burgers-adjoint.py:68-97, :408): lf, alpha = 0.995, Dirichlet
        driver = measure_adjoint_dirichlet(ctx, batch_total=batch_total, n=n, nsteps=min(nsteps, 100), kw=kw, check=check,
                                           flux="lf", alpha=0.995)
    except Exception as exc:  # noqa: BLE001
        import traceback

        traceback.print_exc()
        driver = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
    return {
        "metric": "adjoint gradients/s", "value": batch_total / ((fwd_ms + bwd_ms) * 1e-3), "unit": "gradients/s",
        "n_gpus": world, "steps": nsteps, "scaling": "strong",
        "workload": f"adjoint of a batched Burgers ensemble B={batch_total} x N={n} (rows sharded over {world} rank(s)), "
                    f"{nsteps} fixed-dt SSPRK33 steps, two-level tape (segment {segment}), J = 1/2 sum ||u(T)||^2 "
                    "(BASELINE.json configs[4])",
        "forward_ms": fwd_ms, "reverse_ms": bwd_ms, "ms_per_step": (fwd_ms + bwd_ms) / nsteps,
        "tape_states": tape_states, "segment": segment, "reverse_mode": mode,
        "adjoint_cell_updates_per_s": adj_rate, "forward_cell_updates_per_s": cells * nsteps / (fwd_ms * 1e-3),
        "grad_finite": grad_finite, "parity": parity,
        "roofline": roof,
        "roofline_hbm": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_cell_step": ADJ_ALGO_BYTES,
                         "note": "reverse sweep only, 144 B per cell-step (SURVEY.md 8d per-stage streaming design); the "
                                 "segment recompute of the two-level tape is extra work inside reverse_ms, not extra credit"},
        "gpu_launches": launches,
        "dirichlet": dirichlet,
        "reference_driver_scheme": driver,
    }


# }}}


def run_ours(args: argparse.Namespace) -> None:
    ctx = Ctx()
    line = measure_ensemble(ctx, args)
    skip = set(filter(None, args.skip.split(",")))
    subs: dict[str, dict] = {}

    def sub(name: str, fn) -> None:
        if name in skip:
            return
        try:
            subs[name] = fn()
        except Exception as exc:  # noqa: BLE001  (a failing sub-record must not take the headline with it)
            import traceback

            traceback.print_exc()
            subs[name] = {"error": f"{type(exc).__name__}: {exc}"}
            ctx.torch.cuda.empty_cache()

    if ctx.world > 1:
        sub("weak", lambda: measure_weak(ctx, args))
    sub("slab", lambda: measure_slab(ctx, args, n_global=args.cells or (1 << 30), transport=args.transport,
                                     steps=max(args.steps, 10), warmup=args.warmup))
    sub("adjoint", lambda: measure_adjoint(ctx, args, batch_total=4096, n=8192, nsteps=args.adjoint_steps))
    sub("schemes", lambda: measure_schemes(ctx, args))
    if ctx.rank == 0:
        line.update(subs)
        if not args.no_cpu_baseline:
            # the reference arm's protocol (`--impl reference`: K steps after W warm-up, one step per call) on a shorter
            # run; 3 steps after 1 warm-up read 20-30 % low (first-touch page faults, OpenMP threads spinning up)
            res = cpu_port_steps(10, 3)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    ctx.close()


def run_slab(args: argparse.Namespace) -> None:
    """BASELINE.json configs[3] alone (profiling runs): `--workload slab [--transport ..] [--cells ..]`."""
    ctx = Ctx()
    if ctx.world == 1 and args.transport == "nccl":
        ctx.dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)
    rec = measure_slab(ctx, args, n_global=args.cells or (1 << 30), transport=args.transport, steps=args.steps,
                       warmup=args.warmup, graph=args.graph, check=not args.no_check)
    if ctx.rank == 0:
        rec.update({"higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": rec["workload"]}, "roofline": rec["roofline_hbm"]})
        emit(rec)
    ctx.close()


def run_slab_adjoint(args: argparse.Namespace) -> None:
    """The discrete adjoint of BASELINE configs[3]'s solve on slabs (PeerSlabAdjoint): `--workload slab-adjoint
    [--cells n] [--steps K]`: forward sweep onto per-slab tapes + gathered reverse sweep, one exchange of 16 cells per
    side per step in either sweep."""
    ctx = Ctx()
    torch = ctx.torch
    from pyshocks_b200.distributed import PeerSlabAdjoint

    n_global = args.cells or (1 << 27)
    nsteps = args.steps
    h = (DOMAIN[1] - DOMAIN[0]) / n_global
    sa = PeerSlabAdjoint(n_global=n_global, rank=ctx.rank, world=ctx.world, dx=h, nsteps=nsteps, device=ctx.dev)
    if ctx.world > 1:
        sa.connect()
    else:
        sa.attach_local([sa])
    i = torch.arange(sa.first, sa.first + sa.n_local, device=ctx.dev, dtype=torch.float64)
    u0 = 0.5 + torch.sin(2.0 * np.pi * (i + 0.5) / n_global)
    del i
    dt = torch.full((1,), CFL * h / 1.5, dtype=torch.float64, device=ctx.dev)
    sa.gradient_half_l2(u0, dt)  # warm-up (every kernel loaded, tape pages touched)
    ctx.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    sa.forward_begin(u0)
    for m in range(nsteps):
        sa.forward_step(m, dt)
    ev[1].record()
    sa.backward_begin(sa.interior(nsteps))
    for m in range(nsteps - 1, -1, -1):
        sa.backward_step(m, dt)
    ev[2].record()
    ctx.barrier()
    sa.check()
    fwd_ms, bwd_ms = ctx.max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])])
    finite = bool(torch.isfinite(sa.interior(sa._cur)).all())
    sa.close()
    if ctx.rank == 0:
        emit({
            "metric": "adjoint cell-updates/s (reverse sweep on slabs)", "value": n_global * nsteps / (bwd_ms * 1e-3),
            "unit": "cell-updates/s", "n_gpus": ctx.world, "steps": nsteps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"discrete adjoint of ONE periodic Burgers grid N={n_global}, slab-decomposed over "
                                   f"{ctx.world} rank(s), {nsteps} fixed-dt steps, J = 1/2 ||u(T)||^2: forward sweep onto per-slab "
                                   "tapes, gathered reverse sweep (psk_ssprk33_step_adjoint on slabs with 16 ghost cells)"},
            "forward_ms": fwd_ms, "reverse_ms": bwd_ms, "forward_cell_updates_per_s": n_global * nsteps / (fwd_ms * 1e-3),
            "halo_exchanges_per_step": 1, "halo_cells_per_side": 16, "grad_finite": finite,
            "gpu_launches": 3 * 2 * nsteps + 2,
        })
    ctx.close()


def run_adjoint(args: argparse.Namespace) -> None:
    """BASELINE.json configs[4] alone: `--workload adjoint [--steps K] [--cells n] [--batch B]`."""
    ctx = Ctx()
    rec = measure_adjoint(ctx, args, batch_total=(args.batch if args.batch != BATCH else 4096),
                          n=args.cells or 8192, nsteps=args.steps, check=not args.no_check)
    if ctx.rank == 0:
        rec.update({"higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": rec["workload"]}, "roofline": rec.get("roofline") or rec["roofline_hbm"]})
        emit(rec)
    ctx.close()


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    print(json.dumps(line), file=out, flush=True)


def main() -> None:
    global _JSON_OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=("all", "slab", "adjoint", "slab-adjoint"), default="all",
                    help="all = the headline ensemble + the slab / adjoint sub-records (default); slab / adjoint = "
                         "BASELINE configs 4 and 5 alone")
    ap.add_argument("--transport", choices=("p2p", "p2p-overlap", "p2p-serial", "p2p-step", "p2p-step-fused", "nccl"),
                    default="p2p-step-fused", help="ghost-cell exchange of the slab workload")
    ap.add_argument("--graph", action="store_true", help="slab workload, fused transport: replay a CUDA graph of two steps")
    ap.add_argument("--cells", type=int, default=0, help="override the cell count of the slab / adjoint workloads")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--adjoint-steps", type=int, default=1000, help="steps of the adjoint sub-record (configs[4]: 1000)")
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--batch", type=int, default=BATCH, help="rows of the WHOLE ensemble (default: the named config)")
    ap.add_argument("--skip", default="", help="comma list of sub-records to skip: weak,slab,adjoint,schemes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="slab / adjoint workloads: skip the oracle checks")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", __file__, *sys.argv[1:]]
        raise SystemExit(subprocess.call(cmd))
    # stdout carries the JSON line and nothing else: anything a library prints to fd 1 (NCCL's
    # version banner at communicator creation) goes to stderr instead
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "slab":
        run_slab(args)
    elif args.workload == "adjoint":
        run_adjoint(args)
    elif args.workload == "slab-adjoint":
        run_slab_adjoint(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
