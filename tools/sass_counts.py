"""Static FP64-pipe instruction counts of the hot kernels in libpsk.so -> profiles/sass_counts.json.

    python tools/sass_counts.py            # regenerate (no GPU needed: cuobjdump -sass)

bench.py reads the file to turn a measured rate into an occupancy of the FP64 pipe
(`roofline.frac`): executed FP64-pipe warp instructions per cell-update = static count per lane
of the straight-line kernel / cells a lane emits.  `pyshocks_b200/_build.py` re-runs this after
every build, and the file records the size + mtime-independent SHA-1 of the library it was taken
from, so a stale count is detectable (bench.py compares the hash).

Only straight-line kernels are listed with `per_lane` counts; kernels with data-independent
loops (the fused reverse step) give their loop structure in `loops` and the executed count is
derived in bench.py from it; ncu's executed-instruction counters (profiles/*ncu_summary*) are the
cross-check.
"""
from __future__ import annotations

import collections
import hashlib
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))

from sass_mix import LIB, kernels, registers  # noqa: E402

OUT = ROOT / "profiles" / "sass_counts.json"

# demangled-name prefixes of the kernels bench.py reports on
WANTED = {
    "step_fused": "void psk::step_warp_fused_kernel<6, 0, false, 32, 16, false, 0, 0, false>",
    "step_fused_stages": "void psk::step_warp_fused_kernel<6, 0, false, 32, 16, true, 0, 0, false>",
    "stage1": "void psk::stage_warp_fast_share_kernel<0, 0, 1, false, 0>",
    "stage2": "void psk::stage_warp_fast_share_kernel<0, 0, 2, false, 0>",
    "stage3": "void psk::stage_warp_fast_share_kernel<0, 0, 3, false, 0>",
    "adjoint_lean": "void psk::adjoint_lean_kernel<4, 3>",
    "reverse_step": "void psk::reverse_step_kernel<",
}


def lib_sha1(lib: pathlib.Path = LIB) -> str:
    return hashlib.sha1(lib.read_bytes()).hexdigest()


def main() -> None:
    ks = kernels()
    regs = registers()
    names = list(ks)
    demangled = subprocess.run(["c++filt", *names], capture_output=True, text=True, check=True).stdout.splitlines()
    out = {"library_sha1": lib_sha1(), "how": "cuobjdump -sass, static count per lane (tools/sass_counts.py)", "kernels": {}}
    for key, prefix in WANTED.items():
        for name, dem in zip(names, demangled):
            if not dem.startswith(prefix):
                continue
            c = collections.Counter(op.split(".")[0] for op in ks[name])
            out["kernels"].setdefault(key, []).append({
                "name": dem.split("(")[0],
                "fp64": c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"],
                "dfma": c["DFMA"], "dmul": c["DMUL"], "dadd": c["DADD"], "dsetp": c["DSETP"],
                "mufu": c["MUFU"], "shfl": c["SHFL"], "lds": c["LDS"], "sts": c["STS"], "ldg": c["LDG"], "stg": c["STG"],
                "ldl": c["LDL"], "stl": c["STL"], "bra": c["BRA"], "total": len(ks[name]),
                "resources": regs.get(name, ""),
            })
    missing = [key for key in WANTED if key not in out["kernels"]]
    if missing:  # a template signature changed: bench.py would silently fall back to other counts
        raise SystemExit(f"tools/sass_counts.py: no kernel in {LIB.name} matches {[WANTED[k] for k in missing]}")
    OUT.write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {OUT} ({sum(len(v) for v in out['kernels'].values())} kernels)")


if __name__ == "__main__":
    main()
