"""DRAM traffic and executed FP64 instructions per cell of a profiled launch -> profiles/traffic.json (bench.py reads it).

    python tools/traffic_from_ncu.py <key> <report.ncu-rep> <cells in the profiled launch> [source note]

key: step_kernel | reverse_kernel.  Reads `ncu -i <report> --page raw --csv` (first kernel of the report):
dram__bytes_read.sum + dram__bytes_write.sum, the executed DFMA / DMUL / DADD thread instructions and the FP64
pipe utilisation."""
from __future__ import annotations

import csv
import io
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles" / "traffic.json"


def num(cell: str, unit: str = "") -> float:
    v = float(cell.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    return v * scale


def main() -> None:
    key, rep, cells = sys.argv[1], sys.argv[2], float(sys.argv[3])
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = {k: (v, u) for k, v, u in zip(hdr, r, units)}
    dram = num(*d["dram__bytes_read.sum"]) + num(*d["dram__bytes_write.sum"])
    cyc = num(d["sm__cycles_elapsed.avg"][0])
    fp64 = sum(num(d[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"][0]) for op in ("dfma", "dmul", "dadd"))
    entry = {
        "source": f"{pathlib.Path(rep).name} (ncu --set full --clock-control none, kernel {d['Kernel Name'][0][:60]}) {note}".strip(),
        "cells_in_profiled_launch": cells,
        "dram_bytes_per_cell_measured": dram / cells,
        "fp64_thread_instructions_per_cell_executed": fp64 * cyc / cells,
        "fp64_pipe_pct": num(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
        "duration_ms": num(d["gpu__time_duration.sum"][0]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(d["gpu__time_duration.sum"][1], 1.0),
        "registers": int(num(d["launch__registers_per_thread"][0])),
    }
    cur = json.loads(OUT.read_text()) if OUT.exists() else {}
    cur[key] = entry
    if key == "step_kernel":  # keys bench.py has read since round 1
        cur["step_kernel"]["dram_bytes_per_cell_update_measured"] = entry["dram_bytes_per_cell_measured"]
        cur["step_kernel"]["algorithmic_bytes_per_cell_update"] = 64
    OUT.write_text(json.dumps(cur, indent=1) + "\n")
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
