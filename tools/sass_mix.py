"""Static SASS instruction mix of the kernels in libpsk.so (no GPU needed).

    python tools/sass_mix.py <substring of the mangled kernel name> [...]

Prints, per matching kernel: total instructions, FP64-pipe instructions (DFMA / DMUL / DADD /
DSETP), MUFU, SHFL, LDG / STG, local-memory traffic (LDL / STL = spills) and registers.  The
counts are static (every instruction once, both sides of every branch), which for the
straight-line stage kernels is what one warp executes on the vectorised path plus the
(rarely taken) row-edge path.
"""
from __future__ import annotations

import collections
import pathlib
import re
import subprocess
import sys

LIB = pathlib.Path(__file__).resolve().parents[1] / "pyshocks_b200" / "csrc" / "libpsk.so"


def kernels(lib: pathlib.Path = LIB) -> dict[str, list[str]]:
    text = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    out: dict[str, list[str]] = {}
    name = None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name is not None:
            out[name].append(m.group(2))
    return out


def registers(lib: pathlib.Path = LIB) -> dict[str, str]:
    text = subprocess.run(["cuobjdump", "-res-usage", str(lib)], capture_output=True, text=True, check=True).stdout
    regs, name = {}, None
    for line in text.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?LOCAL:(\d+)", line)
        if m and name:
            regs[name] = f"regs {m.group(1)} local {m.group(2)} B"
    return regs


def main() -> None:
    pats = sys.argv[1:] or ["stage_warp_fast"]
    ks = kernels()
    regs = registers()
    for name, ops in sorted(ks.items()):
        if not any(p in name for p in pats):
            continue
        c = collections.Counter(op.split(".")[0] for op in ops)
        fp64 = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"]
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(f"{demangled}\n   total {len(ops)}  fp64 {fp64} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DSETP {c['DSETP']})"
              f"  MUFU {c['MUFU']}  SHFL {c['SHFL']}  LDG {c['LDG']}  STG {c['STG']}  LDL {c['LDL']}  STL {c['STL']}"
              f"  BRA {c['BRA']}  {regs.get(name, '')}")


if __name__ == "__main__":
    main()
