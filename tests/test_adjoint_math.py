"""The lean adjoint arithmetic of the CUDA kernel (pyshocks_b200/csrc/psk_adjoint_math.cuh),
compiled for the host with g++, against the reference's reconstruction (oracle restatement of
reconstruction.py:351-377 / weno.py:114-256) and its reverse-mode derivative (torch twin):
no GPU needed."""

from __future__ import annotations

import ctypes as ct
import pathlib
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import pyshocks_oracle as po
from oracle import torch_twin as tw

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory: pytest.TempPathFactory) -> ct.CDLL:
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("host") / "libadjmath.so"
    subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(out),
                    str(ROOT / "tests" / "host" / "adjoint_math_host.cpp")], check=True)
    lib = ct.CDLL(str(out))
    dp = ct.POINTER(ct.c_double)
    lib.lean_reconstruct_vjp.argtypes = [ct.c_int, dp, ct.c_double, dp, dp, dp, dp, dp]
    lib.lean_reconstruct_vjp.restype = None
    return lib


def _run(lib: ct.CDLL, f: np.ndarray, eps: float, gl: np.ndarray, gr: np.ndarray):
    n = f.size
    fl, fr, d = np.zeros(n), np.zeros(n), np.zeros(n)
    p = lambda a: a.ctypes.data_as(ct.POINTER(ct.c_double))  # noqa: E731
    lib.lean_reconstruct_vjp(n, p(f), eps, p(gl), p(gr), p(fl), p(fr), p(d))
    return fl, fr, d


def _fields(n: int) -> dict[str, np.ndarray]:
    x = (np.arange(n) + 0.5) / n
    rng = np.random.default_rng(5)
    return {
        "smooth": 0.3 + np.sin(2 * np.pi * x) + 0.2 * np.cos(6 * np.pi * x + 0.4),
        "tophat": np.where((x > 0.3) & (x < 0.7), 1.0, 0.0) + 1e-3 * np.sin(40 * x),
        "noise": rng.standard_normal(n),
        "kink": np.abs(x - 0.47) * 3.0,
    }


@pytest.mark.parametrize("name", ["smooth", "tophat", "noise", "kink"])
def test_lean_forward_values_match_reference_reconstruction(host_lib: ct.CDLL, name: str) -> None:
    n, eps = 64, 1.0e-12
    f = _fields(n)[name]
    z = np.zeros(n)
    fl, fr, _ = _run(host_lib, f, eps, z, z)
    rl, rr = po.reconstruct(po.make_reconstruction("wenojs53", eps), f)
    scale = np.abs(f).max()
    s = slice(2, n - 2)
    assert np.abs(fl[s] - rl[s]).max() <= 2e-14 * scale
    assert np.abs(fr[s] - rr[s]).max() <= 2e-14 * scale


@pytest.mark.parametrize("name,tol", [("smooth", 1e-12), ("noise", 1e-11), ("kink", 1e-10), ("tophat", 1e-9)])
def test_lean_vjp_matches_reverse_mode_of_reference(host_lib: ct.CDLL, name: str, tol: float) -> None:
    n, eps = 64, 1.0e-12
    f = _fields(n)[name]
    rng = np.random.default_rng(11)
    gl, gr = rng.standard_normal(n), rng.standard_normal(n)
    gl[:2] = gl[-2:] = gr[:2] = gr[-2:] = 0.0  # cells whose stencil touches the zero padding
    _, _, d = _run(host_lib, f, eps, gl, gr)
    ft = torch.tensor(f, dtype=torch.float64, requires_grad=True)
    tl, tr = tw.reconstruct(po.make_reconstruction("wenojs53", eps), ft)
    obj = (tl * torch.tensor(gl)).sum() + (tr * torch.tensor(gr)).sum()
    (ref,) = torch.autograd.grad(obj, ft)
    ref = ref.numpy()
    assert np.abs(d - ref).max() <= tol * max(1.0, np.abs(ref).max())
