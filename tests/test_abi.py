"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/psk.h declares, the ctypes mirror of psk_desc has the C layout, and the
product never reaches into the oracle."""

from __future__ import annotations

import ctypes as ct
import pathlib
import re
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "psk.h").read_text()


def declared_functions() -> list[str]:
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(psk_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_every_declared_symbol() -> None:
    from pyshocks_b200 import _lib

    names = declared_functions()
    assert len(names) >= 12
    lib = ct.CDLL(str(_lib.LIB_PATH))
    for name in names:
        assert hasattr(lib, name), f"libpsk.so does not export {name}"
    assert sorted(_lib.EXPORTS) == names
    assert lib.psk_version() == int(re.search(r"#define PSK_VERSION (\d+)", HEADER).group(1))


def test_enum_values_match_header() -> None:
    from pyshocks_b200 import _lib as L

    def enum(name: str) -> dict[str, int]:
        body = re.search(r"enum " + name + r"\s*\{(.*?)\};", HEADER, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return {k: int(v) for k, v in re.findall(r"(PSK_[A-Z_0-9]+)\s*=\s*(\d+)", body)}

    assert enum("psk_equation") == {"PSK_EQ_BURGERS": L.EQ_BURGERS, "PSK_EQ_ADVECTION": L.EQ_ADVECTION, "PSK_EQ_CONTINUITY": L.EQ_CONTINUITY}
    assert enum("psk_flux") == {"PSK_FLUX_RUSANOV": L.FLUX_RUSANOV, "PSK_FLUX_LAX_FRIEDRICHS": L.FLUX_LAX_FRIEDRICHS,
                                "PSK_FLUX_UPWIND": L.FLUX_UPWIND, "PSK_FLUX_ENGQUIST_OSHER": L.FLUX_ENGQUIST_OSHER,
                                "PSK_FLUX_ESWENO": L.FLUX_ESWENO}
    assert enum("psk_rec") == {"PSK_REC_CONSTANT": L.REC_CONSTANT, "PSK_REC_WENOJS32": L.REC_WENOJS32, "PSK_REC_WENOJS53": L.REC_WENOJS53,
                               "PSK_REC_ESWENO32": L.REC_ESWENO32}
    assert enum("psk_bc") == {"PSK_BC_PERIODIC": L.BC_PERIODIC, "PSK_BC_DIRICHLET": L.BC_DIRICHLET,
                              "PSK_BC_NEUMANN": L.BC_NEUMANN, "PSK_BC_NONE": L.BC_NONE}
    assert enum("psk_math") == {"PSK_MATH_FAST": L.MATH_FAST, "PSK_MATH_STRICT": L.MATH_STRICT}
    assert enum("psk_status")["PSK_E_NONFINITE"] == L.E_NONFINITE


def test_desc_layout_matches_c(tmp_path: pathlib.Path) -> None:
    from pyshocks_b200._lib import PskDesc
    from oracle.c_oracle import Desc

    src = tmp_path / "layout.c"
    fields = [f[0] for f in PskDesc._fields_]
    prints = "\n".join(f'  printf("{f} %zu\\n", offsetof(psk_desc, {f}));' for f in fields)
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "psk.h"\nint main(void) {\n'
        '  printf("sizeof %zu\\n", sizeof(psk_desc));\n' + prints + "\n  return 0;\n}\n"
    )
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for struct in (PskDesc, Desc):
        assert ct.sizeof(struct) == int(out["sizeof"])
        for f in fields:
            assert getattr(struct, f).offset == int(out[f]), f


def test_invalid_arguments_are_rejected_without_a_gpu() -> None:
    """argument validation happens on the host, before any CUDA call"""
    from pyshocks_b200 import _lib as L

    d = L.PskDesc()
    lib = L.lib()
    assert lib.psk_apply_operator(ct.byref(d), None, None, None, None) == L.E_INVALID  # n = 0
    d.n, d.g, d.batch, d.ld, d.dx = 16, 1, 1, 18, 0.1
    d.rec = L.REC_WENOJS53
    assert lib.psk_apply_operator(ct.byref(d), None, None, None, None) == L.E_INVALID  # g < stencil width
    d.g, d.ld = 3, 22
    d.equation, d.flux = L.EQ_ADVECTION, L.FLUX_RUSANOV
    assert lib.psk_apply_operator(ct.byref(d), None, None, None, None) == L.E_UNSUPPORTED
    d.equation = 7
    assert lib.psk_apply_operator(ct.byref(d), None, None, None, None) == L.E_UNSUPPORTED
    assert lib.psk_status_string(L.E_UNSUPPORTED).decode().startswith("outside")
    assert lib.psk_set_stage_variant(5) == L.E_INVALID
    # ESWENO32: the Burgers scheme needs its own reconstruction (burgers/schemes.py:214-216)
    d.equation, d.flux, d.rec = L.EQ_BURGERS, L.FLUX_ESWENO, L.REC_WENOJS32
    assert lib.psk_apply_operator(ct.byref(d), None, None, None, None) == L.E_INVALID
    # the fused peer-memory stage: argument checks and the scope of the fused kernel
    link = L.PskHaloLink()
    d.flux, d.rec, d.bc, d.math = L.FLUX_RUSANOV, L.REC_WENOJS53, L.BC_NONE, L.MATH_FAST
    args = (ct.byref(d), 1, None, 1 << 12, 1 << 13, 1 << 14, None)
    assert lib.psk_ssprk33_stage_p2p(*args, None, None) == L.E_INVALID  # no link
    assert lib.psk_ssprk33_stage_p2p(*args, ct.byref(link), None) == L.E_INVALID  # timeout_ns <= 0
    link.timeout_ns = 1
    link.peer_lo = 1 << 15
    assert lib.psk_ssprk33_stage_p2p(*args, ct.byref(link), None) == L.E_INVALID  # peer slot without a flag
    link.peer_lo = None
    d.math = L.MATH_STRICT
    assert lib.psk_ssprk33_stage_p2p(*args, ct.byref(link), None) == L.E_UNSUPPORTED  # FAST math only
    d.math, d.n, d.ld = L.MATH_FAST, 18, 24
    assert lib.psk_ssprk33_stage_p2p(*args, ct.byref(link), None) == L.E_UNSUPPORTED  # n % 4 != 0
    assert lib.psk_halo_push(None, None, None, None, 0, None, None, 1, None) == L.E_INVALID
    assert lib.psk_halo_wait(None, None, 1, 10, None, None) == L.E_INVALID


def test_product_never_touches_the_oracle() -> None:
    """no import, include, dlopen or subprocess of anything under oracle/ from the product"""
    bad = re.compile(r"(^|\s)(from|import)\s+oracle\b|libpsk_oracle|oracle[/.]psk_oracle|#include\s+\"[^\"]*oracle|c_oracle|pyshocks_oracle|torch_twin")
    for path in list((ROOT / "pyshocks_b200").rglob("*")) + [ROOT / "include" / "psk.h"]:
        if path.suffix in {".py", ".cu", ".cuh", ".h"}:
            for ln in path.read_text().splitlines():
                code = ln.split("#")[0] if path.suffix == ".py" else ln
                assert not bad.search(code) or "bit-identical to" in ln or ln.lstrip().startswith(("//", "*", "/*")), (path, ln)


def test_no_cpu_fallback_without_the_library(tmp_path: pathlib.Path) -> None:
    """importing the binding with libpsk.so missing fails loudly"""
    code = (
        "import sys, pathlib; sys.path.insert(0, %r)\n"
        "import pyshocks_b200._lib as L\n" % str(ROOT)
    )
    import os
    import shutil

    fake = tmp_path / "repo"
    shutil.copytree(ROOT / "pyshocks_b200", fake / "pyshocks_b200", ignore=shutil.ignore_patterns("*.so", "__pycache__"))
    code = "import sys; sys.path.insert(0, %r)\nimport pyshocks_b200\n" % str(fake)
    res = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr


def test_static_instruction_counts_cover_the_kernels_bench_reports_on() -> None:
    """profiles/sass_counts.json (regenerated by every build) must hold the whole-step, stage and reverse kernels:
    bench.py's FP64-pipe roofline uses the whole-step count and would otherwise fall back to the stage kernels'"""
    import json
    import pathlib

    d = json.loads((pathlib.Path(__file__).resolve().parents[1] / "profiles" / "sass_counts.json").read_text())
    for key in ("step_fused", "step_fused_stages", "stage1", "stage2", "stage3", "reverse_step"):
        assert d["kernels"].get(key), key
    assert d["kernels"]["step_fused"][0]["fp64"] > 800
