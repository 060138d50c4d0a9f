"""In-tree build of ``libpsk.so`` (hand-written CUDA for sm_100a) with nvcc.

The shared library is written next to the sources (``pyshocks_b200/csrc/libpsk.so``);
it is git-ignored but travels with the repository snapshot to the GPU box.
"""

from __future__ import annotations

import os
import pathlib
import shutil
import subprocess

CSRC = pathlib.Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libpsk.so"
SOURCES = ("psk_forward.cu", "psk_adjoint.cu", "psk_reverse.cu", "psk_solve.cu", "psk_p2p.cu")
HEADERS = ("psk_common.cuh", "psk_math.cuh", "psk_fast_kernels.cuh", "psk_adjoint_math.cuh", "psk_adjoint_kernels.cuh", "psk_reverse_kernels.cuh", "../../include/psk.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=default",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libpsk.so cannot be built")
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    stamp = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES if (CSRC / s).exists()] + [CSRC / h for h in HEADERS]
    return any(p.stat().st_mtime > stamp for p in deps)


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    """Compile every CUDA source (in parallel) and link them into ``libpsk.so``; returns its path."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    extra = os.environ.get("PSK_NVCC_EXTRA", "").split()
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]

    def compile_one(src: pathlib.Path) -> tuple[int, str, pathlib.Path]:
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, *extra, "-c", "-o", str(obj), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr, obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    log = "".join(r[1] for r in results)
    rc = max(r[0] for r in results)
    if rc == 0:
        cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", str(LIB), *[str(r[2]) for r in results]]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        rc = res.returncode
    (CSRC / "build.log").write_text(log)
    if verbose or rc != 0:
        print(log)
    if rc != 0:
        raise RuntimeError(f"nvcc failed with exit code {rc}; see {CSRC / 'build.log'}")
    # static FP64 instruction counts of the hot kernels (bench.py's pipe-occupancy figure) follow the library
    tool = CSRC.parents[1] / "tools" / "sass_counts.py"
    if tool.exists() and shutil.which("cuobjdump"):
        subprocess.run([os.environ.get("PYTHON", "python"), str(tool)], capture_output=True, text=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=False))
