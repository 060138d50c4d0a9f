"""The reference arm of bench.py (`--impl reference`: the CPU restatement of the reference path timed on the host
cores) on a tiny sample: the JSON line carries the keys of the measurement contract, the OpenMP thread count is set
explicitly (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers), and every rank but 0 prints nothing."""

from __future__ import annotations

import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _run(extra_env: dict[str, str]) -> subprocess.CompletedProcess:
    env = dict(os.environ, PSK_BENCH_CPU_ROWS="64", OMP_NUM_THREADS="1", **extra_env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                           "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_line() -> None:
    res = _run({})
    assert res.returncode == 0, res.stderr
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "WENO5 Burgers cell-updates/s" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 2 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "configs[2]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "64 of 65536 rows" in cb["sample"]
    # the thread count is the affinity core count, not the OMP_NUM_THREADS=1 a torchrun worker inherits
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent() -> None:
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0, res.stderr
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
