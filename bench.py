#!/usr/bin/env python
"""Benchmark of the pyshocks hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric "WENO5 Burgers
cell-updates/s" is quoted on): an ensemble of B = 65536 independent inviscid Burgers
problems x N = 4096 cells, fp64, WENO-JS5 + Rusanov (LLF) + SSPRK33, periodic, random
smooth initial data, one shared fixed dt at CFL 0.4.  One "step" = one full SSPRK33 step
(one launch of the whole-step kernel, psk_ssprk33_step) of every cell of the ensemble.  With N GPUs every rank advances
its own B rows (weak scaling, no collective in the data path).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the
reference path (oracle/psk_oracle.c, all host threads) on a bounded sample of the same
workload; the reference itself is Python-on-JAX and JAX is not installable here.
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


# {{{ reference arm / cpu baseline: the C restatement on the host cores


def cpu_port_throughput(target_seconds: float, steps: int | None = None) -> dict:
    from oracle.c_oracle import COracle

    cores = len(os.sched_getaffinity(0))
    h = (DOMAIN[1] - DOMAIN[0]) / N_CELLS
    # calibrate on a small slice, then size the sample
    rows = max(4 * cores, 32)
    coef = ensemble_coefficients(rows, 20261017)
    u0 = host_initial_condition(coef, N_CELLS, GHOSTS)
    dt = CFL * h / np.abs(u0).max()
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS, g=GHOSTS,
                 batch=rows, dx=h, eps=EPS)
    co.solve_fixed_dt(u0, dt, 1)  # thread pool start-up
    t0 = time.perf_counter()
    co.solve_fixed_dt(u0, dt, 2)
    rate = 2 * rows * N_CELLS / (time.perf_counter() - t0)  # cell-updates/s with all threads
    nsteps = steps if steps is not None else 4
    rows_s = int(min(BATCH, target_seconds * rate / (N_CELLS * nsteps)))
    rows_s = max(cores, (rows_s // cores) * cores)
    coef = ensemble_coefficients(rows_s, 20261017)
    u0 = host_initial_condition(coef, N_CELLS, GHOSTS)
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS, g=GHOSTS,
                 batch=rows_s, dx=h, eps=EPS)
    t0 = time.perf_counter()
    co.solve_fixed_dt(u0, dt, nsteps)
    wall = time.perf_counter() - t0
    return {
        "value": rows_s * N_CELLS * nsteps / wall,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{rows_s} rows x {N_CELLS} cells x {nsteps} SSPRK33 steps of the same ensemble "
                  f"(C restatement oracle/psk_oracle.c, OpenMP over rows, {wall:.1f} s)",
    }


def run_reference(args: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.c_oracle import COracle

    cores = len(os.sched_getaffinity(0))
    h = (DOMAIN[1] - DOMAIN[0]) / N_CELLS
    # a "step" of the reference arm = one SSPRK33 step over a bounded sample of rows
    rows = max(cores * 8, 64)
    coef = ensemble_coefficients(rows, 20261017)
    u0 = host_initial_condition(coef, N_CELLS, GHOSTS)
    dt = CFL * h / np.abs(u0).max()
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS, g=GHOSTS,
                 batch=rows, dx=h, eps=EPS)
    u = u0
    for _ in range(args.warmup):
        u = co.solve_fixed_dt(u, dt, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        u = co.solve_fixed_dt(u, dt, 1)
    wall = time.perf_counter() - t0
    value = rows * N_CELLS * args.steps / wall
    sample = f"{rows} of {BATCH} rows x {N_CELLS} cells per step (bounded sample of the same ensemble)"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus) | {"reference_arm": "CPU restatement of the reference path "
                                               "(oracle/psk_oracle.c); JAX is not installable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# }}}


def measure_fp64_peak(dev) -> dict:
    """DFMA throughput of this GPU (MEASURED_PEAKS.json has no fp64 figure): psk_dfma_probe, best of 5."""
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
    return {"tflops": best / 1e12, "how": "psk_dfma_probe: 16 independent DFMA chains per thread, 2368 CTAs x 256 threads, best of 6"}


# algorithmic flops per cell-update (152 per cell-stage, divisions counted as 1)
ALGO_FLOPS_PER_CELL_UPDATE = 456.0


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": f"batched Burgers ensemble B={BATCH} x N={N_CELLS} cells per GPU, fp64, WENO-JS5 + Rusanov(LLF) "
                    f"+ SSPRK33, periodic, fixed dt at CFL {CFL} (BASELINE.json configs[2])",
        "batch_per_gpu": BATCH, "cells": N_CELLS, "ghosts": GHOSTS,
        "sharding": f"ensemble rows block-partitioned, {n_gpus} rank(s), no data-path collective",
        "l2": "inputs (2.15 GB per state array) are larger than the 126 MB L2; no flush needed",
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
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

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


def run_ours(args: argparse.Namespace) -> None:
    import torch
    import torch.distributed as dist

    from pyshocks_b200.ensemble import EnsembleSolver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier() -> None:
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    batch = args.batch
    h = (DOMAIN[1] - DOMAIN[0]) / N_CELLS
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=N_CELLS,
                            g=GHOSTS, dx=h, eps=EPS, batch=batch, math="fast", device=dev)
    nx = solver.nx

    # synthetic initial data of the named shape, evaluated on the device from host-drawn coefficients
    coef = torch.from_numpy(ensemble_coefficients(batch, 20261017 + rank)).to(dev)
    xhat = ((torch.arange(nx, device=dev, dtype=torch.float64) - GHOSTS + 0.5) / N_CELLS)[None, :]
    u0 = coef[:, :1].repeat(1, nx)
    for k in range(4):
        u0 += coef[:, 1 + k : 2 + k] * torch.sin(2.0 * np.pi * (k + 1) * xhat + coef[:, 5 + k : 6 + k])
    umax = u0.abs().max()
    if world > 1:
        dist.all_reduce(umax, op=dist.ReduceOp.MAX)
    dt = torch.full((1,), CFL * h / float(umax), dtype=torch.float64, device=dev)
    solver.load(u0)

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps
    solver.solve_fixed_dt(None, dt, args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = solver.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    solver.solve_fixed_dt(None, dt, args.steps)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    launches = solver.launches - launches0
    clocks = sampler.stop() if sampler is not None else None
    finite = bool(torch.isfinite(solver.u).all())

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    host_in = torch.empty((batch, nx), dtype=torch.float64, pin_memory=True)
    host_out = torch.empty((batch, nx), dtype=torch.float64, pin_memory=True)
    host_in.copy_(u0)
    del u0
    barrier()
    solver.solve_fixed_dt_host(host_in, host_out, dt, 1)  # warm-up of the copy streams
    barrier()
    t0 = time.perf_counter()
    solver.solve_fixed_dt_host(host_in, host_out, dt, args.steps)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)
    bytes_state = batch * nx * 8

    if rank == 0:
        cells_per_step = batch * N_CELLS * world
        value = cells_per_step * args.steps / (ms_total * 1e-3)
        peak, peak_src = measured_peaks()
        per_gpu = value / world
        achieved = per_gpu * ALGO_BYTES_PER_CELL_UPDATE / 1e9
        per_step = launches / max(args.steps, 1)  # 1: whole-step kernel, 3: one launch per stage
        whole = per_step < 2
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            tj = json.loads(tf.read_text())
            traffic = tj.get("step_kernel_dram_bytes_per_launch" if whole else "stage_kernel_dram_bytes_per_launch")
        fp64 = measure_fp64_peak(dev)
        # whole-step kernel: 922 per lane / (172 emitted cells / 32 lanes); stage kernels: (202 + 210 + 210) / (120 / 32)
        fp64_per_update = 922 * 32 / 172 if whole else 622 * 32 / 120
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world) | {"batch_per_gpu": batch, "finite": finite},
            "roofline_fp64": {
                "bound": "fp64", "achieved": per_gpu * ALGO_FLOPS_PER_CELL_UPDATE / 1e12, "peak": fp64["tflops"],
                "unit": "TFLOP/s", "frac": per_gpu * ALGO_FLOPS_PER_CELL_UPDATE / 1e12 / fp64["tflops"],
                "peak_source": fp64["how"],
                # the same bound in EXECUTED instructions: FP64-pipe warp instructions per cell-update (static
                # SASS count of the kernel over the cells a warp emits, tools/sass_mix.py) against
                # 64 lanes/clk/SM x 148 SMs at the SM clock sampled during the timed region
                "executed_fp64_instr_per_cell_update": fp64_per_update,
                "pipe_frac": (per_gpu * fp64_per_update / (148 * 64 * clocks["sm_mhz"] * 1e6)
                              if clocks and clocks.get("sm_mhz") else None),
                "note": "456 algorithmic flop per cell-update (SURVEY.md 8d) against the measured DFMA peak: the "
                        "FMA-fused, re-associated kernel executes far fewer operations than the reference's arithmetic "
                        "counts, so `frac` can exceed 1; `pipe_frac` is the occupancy of the FP64 pipe by the "
                        "instructions actually executed (ncu: 84 %), the binding unit of this path (DESIGN.md 4.0, 8)",
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": ("psk::step_warp_fused_kernel (1 launch per step: the three stages in registers; 64 "
                           "algorithmic bytes per cell-update, 16 B of them actually moved)") if whole else
                          "psk::stage_warp_fast_share_kernel (3 launches per step; 64 algorithmic bytes per cell-update)",
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL_UPDATE / per_step * batch * N_CELLS,
                "avg_launch_ms": ms_total / max(launches, 1),
            },
            "e2e": {
                "value": cells_per_step * args.steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": bytes_state / args.steps, "d2h_bytes_per_step": bytes_state / args.steps,
                "call": f"EnsembleSolver.solve_fixed_dt_host(pinned host in, pinned host out, dt, K): "
                        f"{max(1, min(64, batch // 1024))} row blocks, upload / K steps / download pipelined over 4 streams",
                "seconds": e2e_s,
            },
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_throughput(target_seconds=15.0)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_slab(args: argparse.Namespace) -> None:
    """BASELINE.json configs[3]: ONE periodic Burgers grid of N = 2^30 cells, slab-decomposed over
    the ranks with ring halo exchange (3 cells per side per stage), fixed dt at CFL 0.4."""
    import torch
    import torch.distributed as dist

    from pyshocks_b200.distributed import DistRing, PeerSlabSolver, SlabSolver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)
    n_global = args.cells if args.cells else (1 << 30)
    h = (DOMAIN[1] - DOMAIN[0]) / n_global
    if args.transport == "nccl":
        slab = SlabSolver(n_global=n_global, ring=DistRing(), dx=h, device=dev)
        launches_per_step = 3
    else:
        whole = args.transport == "p2p-step"
        slab = PeerSlabSolver(n_global=n_global, rank=rank, world=world, dx=h, device=dev, whole_step=whole,
                              overlap=(args.transport == "p2p-overlap"), fused=(args.transport == "p2p"))
        slab.connect()
        # per stage: fused = 1 launch; else wait, (2 edge +) 1 stage kernel, push; whole step: wait, step, push
        launches_per_step = 3 if (slab.fused or whole) else (15 if slab.split else 9)
    i = torch.arange(slab.first, slab.first + slab.n_local, device=dev, dtype=torch.float64)
    slab.load_interior(0.5 + torch.sin(2.0 * np.pi * (i + 0.5) / n_global))
    del i
    dt = torch.full((1,), CFL * h / 1.5, dtype=torch.float64, device=dev)

    def barrier() -> None:
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    kw = {"graph": True} if (args.transport == "p2p" and args.graph) else {}
    slab.solve_fixed_dt(dt, max(args.warmup, 6 if kw else 0), **kw)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    slab.solve_fixed_dt(dt, args.steps, **kw)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if args.transport != "nccl":
        slab.check()  # a ghost-cell wait that timed out would have produced garbage silently
    finite = bool(torch.isfinite(slab.interior()).all().item())
    if rank == 0:
        value = n_global * args.steps / (float(ms) * 1e-3)
        peak, peak_src = measured_peaks()
        achieved = value / world * ALGO_BYTES_PER_CELL_UPDATE / 1e9
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(ms) / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"single periodic Burgers grid N={n_global} cells, slab-decomposed over {world} rank(s), "
                                   + ("ring halo exchange 9 cells/side/step" if args.transport == "p2p-step" else
                                      "ring halo exchange 3 cells/side/stage") + " (BASELINE.json configs[3])",
                       "cells_per_gpu": slab.n_local,
                       "halo_exchanges_per_step": 1 if args.transport == "p2p-step" else 3, "finite": finite,
                       "cuda_graph": bool(kw),
                       "transport": {"p2p": "exchange fused into the stage kernel: edge warps spin on local epoch flags, "
                                            "edge lanes store into the neighbours' ghost slots over NVLink (1 launch per stage)",
                                     "p2p-overlap": "NVLink peer stores + epoch flags, slab edges on a high-priority stream "
                                                    "overlapped with the interior",
                                     "p2p-serial": "NVLink peer stores + epoch flags, no overlap",
                                     "p2p-step": "whole SSPRK33 step in one launch on a slab with 9 ghost cells; one "
                                                 "exchange per step (NVLink peer stores + epoch flags, no overlap)",
                                     "nccl": "NCCL send/recv pairs per stage"}[args.transport]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src},
            "gpu_launches": launches_per_step * args.steps,
        })
    if args.transport != "nccl":
        slab.close()
    dist.destroy_process_group()


def run_adjoint(args: argparse.Namespace) -> None:
    """BASELINE.json configs[4]: B = 4096 x N = 8192 ensemble, K fixed-dt steps forward with a
    two-level device tape, reverse sweep for J = 1/2 sum ||u(T)||^2; rows sharded over the ranks."""
    import torch
    import torch.distributed as dist

    from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch, n = (args.batch if args.batch != BATCH else 4096), (args.cells if args.cells else 8192)
    h = (DOMAIN[1] - DOMAIN[0]) / n
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=GHOSTS,
                            dx=h, eps=EPS, batch=batch, device=dev)
    coef = torch.from_numpy(ensemble_coefficients(batch, 20261018 + rank)).to(dev)
    xhat = ((torch.arange(solver.nx, device=dev, dtype=torch.float64) - GHOSTS + 0.5) / n)[None, :]
    u0 = coef[:, :1].repeat(1, solver.nx)
    for k in range(4):
        u0 += coef[:, 1 + k : 2 + k] * torch.sin(2.0 * np.pi * (k + 1) * xhat + coef[:, 5 + k : 6 + k])
    dt = CFL * h / float(u0.abs().max())
    nsteps = args.steps
    adj = AdjointEnsemble(solver, nsteps=nsteps, dt=dt)
    small = AdjointEnsemble(solver, nsteps=4, dt=dt, segment=2)
    small.gradient_half_l2(u0)  # warm-up of every kernel
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    if world > 1:
        dist.barrier(device_ids=[local])
    ev[0].record()
    uT = adj.forward(u0)
    ev[1].record()
    pT = adj.lam1
    pT.zero_()
    pT[:, GHOSTS : GHOSTS + n] = uT[:, GHOSTS : GHOSTS + n]
    grad = adj.backward(pT)
    ev[2].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fwd_ms, bwd_ms = float(t[0]), float(t[1])
    if rank == 0:
        peak, peak_src = measured_peaks()
        cells = batch * n * world
        adj_rate = cells * nsteps / (bwd_ms * 1e-3)
        # reverse sweep, per cell-step: recompute k1, k2 (16 + 24 B), three adjoint stages (32 + 40 + 32 B),
        # plus the segment recompute (64 B per forward step, amortised (segment - 1) / segment)
        algo = 144.0 + 64.0 * (adj.segment - 1) / adj.segment
        achieved = adj_rate / world * algo / 1e9
        emit({
            "metric": "adjoint gradients/s", "value": batch * world / ((fwd_ms + bwd_ms) * 1e-3), "unit": "gradients/s",
            "n_gpus": world, "steps": nsteps, "warmup": args.warmup, "ms_per_step": (fwd_ms + bwd_ms) / nsteps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"adjoint of a batched Burgers ensemble B={batch} x N={n} per GPU, {nsteps} fixed-dt SSPRK33 "
                                   f"steps, two-level tape (segment {adj.segment}), J = 1/2 sum ||u(T)||^2 (BASELINE.json configs[4])",
                       "forward_ms": fwd_ms, "reverse_ms": bwd_ms, "tape_states": len(adj.chk) + len(adj.ring),
                       "adjoint_cell_updates_per_s": adj_rate, "forward_cell_updates_per_s": cells * nsteps / (fwd_ms * 1e-3),
                       "grad_finite": bool(torch.isfinite(grad).all())},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_cell_step": algo},
            "gpu_launches": adj.launches,
        })
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    print(json.dumps(line), file=out, flush=True)


def main() -> None:
    global _JSON_OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=("ensemble", "slab", "adjoint"), default="ensemble",
                    help="ensemble = the headline config (default); slab / adjoint = BASELINE configs 4 and 5")
    ap.add_argument("--transport", choices=("p2p", "p2p-overlap", "p2p-serial", "p2p-step", "nccl"), default="p2p",
                    help="ghost-cell exchange of the slab workload")
    ap.add_argument("--graph", action="store_true", help="slab workload, fused transport: replay a CUDA graph of two steps")
    ap.add_argument("--cells", type=int, default=0, help="override the cell count of the slab / adjoint workloads")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--batch", type=int, default=BATCH, help="rows per GPU (default: the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
