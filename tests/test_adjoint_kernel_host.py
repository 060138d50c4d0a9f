"""The lean adjoint stage kernel of the product (pyshocks_b200/csrc/psk_adjoint_kernels.cuh: the
hand-derived transpose of the fused WENO-JS5 + Rusanov stage, what the reference's adjoint_step gets
from jax.jacfwd, timestepping.py:174, :205-206) compiled for the HOST and run under the warp emulation
of tests/host/emu/cuda_runtime.h, against reverse-mode differentiation of the reference arithmetic
(oracle/torch_twin.py).  Pins the transposed stencil, the lane exchange of the cotangents, the row
tails and the ghost-cell spill of the kernel without a GPU."""

from __future__ import annotations

import ctypes as ct
import pathlib
import shutil
import subprocess

import numpy as np
import pytest

from oracle import pyshocks_oracle as po
from oracle import torch_twin as tt

ROOT = pathlib.Path(__file__).resolve().parent.parent
G = 3
EPS = 1.0e-12


@pytest.fixture(scope="module")
def emu(tmp_path_factory: pytest.TempPathFactory) -> ct.CDLL:
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("emu") / "libadjemu.so"
    subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-pthread",
                    "-I", str(ROOT / "tests" / "host" / "emu"), "-o", str(out),
                    str(ROOT / "tests" / "host" / "adjoint_kernel_host.cpp")], check=True)
    lib = ct.CDLL(str(out))
    dp = ct.POINTER(ct.c_double)
    lib.emu_adjoint_lean.argtypes = [ct.c_int] * 4 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, ct.c_int,
                                                      ct.c_double, ct.c_double, dp, ct.c_double, dp, ct.c_double, dp, dp]
    lib.emu_adjoint_lean.restype = ct.c_int
    lib.emu_adjoint_lean_flux.argtypes = [ct.c_int] * 5 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, ct.c_int,
                                                           ct.c_double, ct.c_double, dp, dp, dp, dp]
    lib.emu_adjoint_lean_flux.restype = ct.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ct.POINTER(ct.c_double))


def _state(n: int, kind: str, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = (np.arange(n + 2 * G) - G + 0.5) / n
    u = rng.uniform(-0.3, 0.3) + sum(rng.uniform(0, 1 / k) * np.sin(2 * np.pi * k * x + rng.uniform(0, 6.28))
                                     for k in range(1, 4))
    if kind == "tophat":
        u = u + np.where((x > 0.3) & (x < 0.6), 0.8, 0.0)
    if kind == "plateau":  # several cells hold max |u| exactly (the speed cotangent is shared between them)
        u = np.minimum(u, 0.9 * np.abs(u).max())
    return u


def _run(emu, variant: int, n: int, x: np.ndarray, v: np.ndarray, dt: float, c_v: float, c_g: float,
         acc=None, c_acc=0.0, acc2=None, c_acc2=0.0) -> np.ndarray:
    nx = n + 2 * G
    batch = x.shape[0]
    out = np.full((batch, nx), np.nan)
    gspill = np.zeros((batch, 2 * G))
    dts = np.full(batch, dt)
    rc = emu.emu_adjoint_lean(variant, n, G, batch, nx, 3.0 / n, EPS, _p(x), _p(v), _p(dts), 1, c_v, c_g, _p(acc), c_acc,
                              _p(acc2), c_acc2, _p(gspill), _p(out))
    assert rc == 0
    return out


@pytest.mark.parametrize("variant", [3, 0])
@pytest.mark.parametrize("n,kind,tol", [(250, "smooth", 1e-12), (121, "smooth", 1e-12), (500, "tophat", 1e-9),
                                        (64, "smooth", 1e-12)])
def test_lean_adjoint_stage_is_the_transposed_jacobian(emu, variant: int, n: int, kind: str, tol: float) -> None:
    rng = np.random.default_rng(n)
    batch = 2
    x = np.stack([_state(n, kind, 10 * n + b) for b in range(batch)])
    v = rng.standard_normal((batch, n + 2 * G))
    dt = 0.3 * (3.0 / n)
    out = _run(emu, variant, n, x, v, dt, 1.0, 1.0)
    assert np.isfinite(out).all()
    scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53", EPS))
    grid = po.make_grid(-1.5, 1.5, n, G)
    for b in range(batch):
        ref = v[b] + dt * tt.rhs_vjp(scheme, grid, po.Periodic(), 0.0, x[b], v[b])
        err = np.abs(out[b] - ref).max() / np.abs(ref).max()
        assert err < tol, (b, err)


def test_lean_adjoint_stage_linear_terms(emu) -> None:
    """out = c_acc acc + c_acc2 acc2 + c_v v + c_g dt J^T v (psk_adjoint.cu); the last call of a reverse
    SSPRK33 step has c_v = c_g = 1, c_acc = 1/3, c_acc2 = 3/4, the first c_v = c_g = 2/3 (include/psk.h,
    psk_ssprk33_stage_adjoint)"""
    n, batch = 300, 2
    rng = np.random.default_rng(1)
    x = np.stack([_state(n, "smooth", b) for b in range(batch)])
    v, acc, acc2 = (rng.standard_normal((batch, n + 2 * G)) for _ in range(3))
    dt = 0.25 * (3.0 / n)
    base = _run(emu, 3, n, x, v, dt, 1.0, 1.0)  # v + dt J^T v
    got = _run(emu, 3, n, x, v, dt, 1.0, 1.0, acc=acc, c_acc=1.0 / 3.0, acc2=acc2, c_acc2=0.75)
    ref = acc / 3.0 + 0.75 * acc2 + base
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()
    got = _run(emu, 3, n, x, v, dt, 2.0 / 3.0, 2.0 / 3.0)
    assert np.abs(got - (2.0 / 3.0) * base).max() <= 1e-14 * np.abs(base).max()
    got = _run(emu, 3, n, x, v, dt, 0.25, 1.0, acc=acc, c_acc=-2.0)
    assert np.abs(got - (-2.0 * acc + 0.25 * v + (base - v))).max() <= 1e-14 * np.abs(base).max()


@pytest.mark.parametrize("flux,alpha", [("lf", 1.0), ("lf", 0.995), ("rusanov", 0.995)])
@pytest.mark.parametrize("bc", ["periodic", "dirichlet", "neumann"])
@pytest.mark.parametrize("n,kind,tol", [(250, "smooth", 1e-12), (121, "smooth", 1e-12), (500, "tophat", 1e-9),
                                        (250, "plateau", 1e-9)])
def test_lean_adjoint_stage_with_the_global_speed_and_the_viscosity_of_every_face(emu, flux: str, alpha: float, bc: str,
                                                                                  n: int, kind: str, tol: float) -> None:
    """the Lax-Friedrichs (scalar.py:258-278) and alpha != 1 (scalar.py:231-234) forms of the lean kernel: the speed
    cotangent summed over the faces of a row and handed to the arg-max cells, nu per face; periodic and Dirichlet rows"""
    rng = np.random.default_rng(n + int(1000 * alpha))
    batch = 2
    x = np.stack([_state(n, kind, 10 * n + b) for b in range(batch)])
    v = rng.standard_normal((batch, n + 2 * G))
    dt = 0.3 * (3.0 / n)
    grid = po.make_grid(-1.5, 1.5, n, G)
    scheme = po.Scheme("burgers", flux, po.make_reconstruction("wenojs53", EPS), alpha=alpha)
    nu = None if alpha == 1.0 else (np.diff(grid.x) ** (alpha - 1.0)).copy()
    ghost = rng.uniform(-0.4, 0.4, size=(batch, 2 * G)) if bc == "dirichlet" else None
    slope = rng.uniform(-0.5, 0.5, size=batch)
    if bc == "neumann":  # scalar.py:472-500: ghost = mirror image + side * (x[ifrom] - x[ito]) * g(t)
        gi = np.arange(G)
        dxl = grid.x[2 * G - 1 - gi] - grid.x[gi]
        ir = grid.nx - G + gi
        dxr = grid.x[2 * (grid.nx - G) - 1 - ir] - grid.x[ir]
        ghost = np.stack([np.concatenate([-slope[b] * dxl, slope[b] * dxr]) for b in range(batch)])
    nx = n + 2 * G
    out = np.full((batch, nx), np.nan)
    gspill = np.zeros((batch, 2 * G))
    dts = np.full(batch, dt)
    assert emu.emu_adjoint_lean_flux(1 if flux == "lf" else 0, {"periodic": 0, "dirichlet": 1, "neumann": 2}[bc], n, G, batch, nx, grid.h, EPS,
                                     _p(x), _p(v), _p(dts), 1, 1.0, 1.0, _p(nu), _p(ghost), _p(gspill), _p(out)) == 0
    assert np.isfinite(out).all()
    xg = np.concatenate([grid.x[:G], grid.x[-G:]])
    for b in range(batch):
        if bc == "periodic":
            obc = po.Periodic()
        elif bc == "dirichlet":
            obc = po.Dirichlet(ga=lambda t, xx, b=b: np.interp(xx, xg, ghost[b]))
        else:
            obc = po.Neumann(ga=lambda t, b=b: slope[b])
        ref = v[b] + dt * tt.rhs_vjp(scheme, grid, obc, 0.0, x[b], v[b])
        err = np.abs(out[b] - ref).max() / np.abs(ref).max()
        assert err < tol, (flux, alpha, bc, b, err)
