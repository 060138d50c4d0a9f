"""The fused reverse SSPRK33 step of the product (pyshocks_b200/csrc/psk_reverse_kernels.cuh: recompute
k1, k2 and apply the three transposed stages in one kernel -- what the reference's adjoint_step gets per
step from jax.jacfwd(advance).T @ p, timestepping.py:174, :198-209) compiled for the HOST and run under
the warp emulation of tests/host/emu/cuda_runtime.h.  Pins the lane-private streams, the exchanges at the
run ends, the window overlap and the periodic wrap without a GPU: against reverse-mode differentiation
of the reference arithmetic (oracle/torch_twin.py) and, for the recomputed stages, against the C oracle."""

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
    out = tmp_path_factory.mktemp("emu") / "librevemu.so"
    subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-pthread",
                    "-I", str(ROOT / "tests" / "host" / "emu"), "-o", str(out),
                    str(ROOT / "tests" / "host" / "reverse_kernel_host.cpp")], check=True)
    lib = ct.CDLL(str(out))
    dp = ct.POINTER(ct.c_double)
    lib.emu_reverse_step.argtypes = ([ct.c_int] * 4 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, ct.c_int, dp, dp, dp,
                                      ct.c_int, dp, ct.c_longlong, ct.c_longlong])
    lib.emu_reverse_step.restype = ct.c_int
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
    return u


def _aligned(batch: int, n: int) -> tuple[np.ndarray, int, int]:
    """(storage, col0, ld): rows whose first interior cell is 16-byte aligned (ensemble.row_layout)"""
    col0 = (16 - G) % 16
    ld = ((col0 + n + 2 * G + 15) // 16) * 16
    raw = np.zeros(batch * ld + 2)
    off = 0 if (raw.ctypes.data + 8 * (col0 + G)) % 16 == 0 else 1
    return raw[off : off + batch * ld].reshape(batch, ld), col0, ld


def _run(emu, C: int, n: int, u: np.ndarray, p: np.ndarray, dt: float, stages: bool = False,
         ghost3: np.ndarray | None = None):
    batch = u.shape[0]
    nx = n + 2 * G
    bufs = []
    for src in (u, p, None, None, None):
        a, col0, ld = _aligned(batch, n)
        v = a[:, col0 : col0 + nx]
        if src is not None:
            v[:] = src
        else:
            v[:] = np.nan
        bufs.append(v)
    U, P, OUT, K1, K2 = bufs
    dts = np.full(batch, dt)
    rc = emu.emu_reverse_step(C, n, G, batch, ld, 3.0 / n, EPS, _p(U), _p(P), _p(dts), 1, _p(OUT),
                              _p(K1) if stages else None, _p(K2) if stages else None, 0,
                              _p(ghost3), 0 if ghost3 is None else 2 * G, 0 if ghost3 is None else batch * 2 * G)
    assert rc == 0
    return (OUT, K1, K2) if stages else OUT


@pytest.mark.parametrize("C", [8, 12, 16, 20, 24])
@pytest.mark.parametrize("n,kind,tol", [(1000, "smooth", 3e-12), (300, "smooth", 3e-12), (74, "smooth", 3e-12),
                                        (812, "tophat", 2e-9), (1300, "smooth", 3e-12)])
def test_fused_reverse_step_is_the_transposed_step_jacobian(emu, C: int, n: int, kind: str, tol: float) -> None:
    # tolerance: three chained stages; the worst cells (next to an extremum, beta ~ eps) are the same ones with the
    # same error for every run length C, i.e. round-off of the derivative itself, not of the tiling
    rng = np.random.default_rng(n + C)
    batch = 2
    u = np.stack([_state(n, kind, 10 * n + b) for b in range(batch)])
    p = rng.standard_normal((batch, n + 2 * G))
    p[:, :G] = 0.0
    p[:, n + G :] = 0.0
    dt = 0.3 * (3.0 / n)
    out = _run(emu, C, n, u, p, dt)
    i = slice(G, G + n)
    assert np.isfinite(out[:, i]).all()
    assert np.isnan(out[:, :G]).all() and np.isnan(out[:, n + G :]).all()  # ghost cells are not written
    scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53", EPS))
    grid = po.make_grid(-1.5, 1.5, n, G)
    for b in range(batch):
        ref = tt.step_vjp(scheme, grid, po.Periodic(), dt, 0.0, u[b], p[b])
        err = np.abs(out[b, i] - ref[i]).max() / np.abs(ref[i]).max()
        assert err < tol, (b, err)


@pytest.mark.parametrize("C,n", [(16, 1000), (12, 74), (20, 2000), (20, 1300), (24, 8192)])
def test_recomputed_stage_values_match_the_oracle(emu, C: int, n: int) -> None:
    """k1, k2 of timestepping.py:314-317 as the kernel recomputes them (FAST arithmetic: 1e-13 of the C oracle)"""
    from oracle.c_oracle import COracle

    batch = 2
    u = np.stack([_state(n, "smooth", 7 * n + b) for b in range(batch)])
    p = np.zeros_like(u)
    dt = 0.3 * (3.0 / n)
    _, k1, k2 = _run(emu, C, n, u, p, dt, stages=True)
    co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=G, batch=batch,
                 dx=3.0 / n, eps=EPS)
    L = co.apply_operator(u)
    r1 = u + dt * L
    r2 = 0.75 * u + 0.25 * (r1 + dt * co.apply_operator(r1))
    i = slice(G, G + n)
    assert np.abs(k1[:, i] - r1[:, i]).max() < 1e-13 * np.abs(r1).max()
    assert np.abs(k2[:, i] - r2[:, i]).max() < 1e-13 * np.abs(r2).max()


@pytest.mark.parametrize("C", [12, 16])
@pytest.mark.parametrize("world,n", [(2, 1000), (3, 812)])
def test_fused_reverse_step_on_slabs_with_sixteen_ghost_cells(emu, C: int, world: int, n: int) -> None:
    """boundary kind NONE: a slab's window cells beyond its ends are its stored ghost cells (16 per side, the
    neighbours' edge cells of u and of p'); the slabs together give the periodic row's result -- every slab
    GATHERS its part of the gradient, nothing is sent back (to round-off: the lane boundaries, hence the order
    in which a first difference collects its cotangents, sit elsewhere)"""
    rng = np.random.default_rng(n)
    u = _state(n, "smooth", n)[None, :]
    p = rng.standard_normal((1, n + 2 * G))
    p[:, :G] = 0.0
    p[:, n + G :] = 0.0
    dt = 0.3 * (3.0 / n)
    whole = _run(emu, C, n, u, p, dt)[0, G : G + n]
    g16, first, parts = 16, 0, []
    ui, pi = u[0, G : G + n], p[0, G : G + n]
    for r in range(world):
        nl = 2 * ((n // world + (1 if r < n % world else 0)) // 2)  # even slab lengths
        if r == world - 1:
            nl = n - first
        idx = np.arange(first - g16, first + nl + g16) % n
        col0 = (16 - g16) % 16
        ld = ((col0 + nl + 2 * g16 + 15) // 16) * 16
        bufs = []
        for src in (ui[idx], pi[idx], None):
            raw = np.zeros(ld + 2)
            off = 0 if (raw.ctypes.data + 8 * (col0 + g16)) % 16 == 0 else 1
            v = raw[off : off + ld].reshape(1, ld)[:, col0 : col0 + nl + 2 * g16]
            v[:] = np.nan if src is None else src
            bufs.append(v)
        U, P, OUT = bufs
        dts = np.full(1, dt)
        assert emu.emu_reverse_step(C, nl, g16, 1, ld, 3.0 / n, EPS, _p(U), _p(P), _p(dts), 1, _p(OUT), None, None, 1,
                                    None, 0, 0) == 0
        assert np.isnan(OUT[0, :g16]).all() and np.isnan(OUT[0, g16 + nl :]).all()
        parts.append(OUT[0, g16 : g16 + nl].copy())
        first += nl
    got = np.concatenate(parts)
    assert np.abs(got - whole).max() < 1e-13 * np.abs(whole).max()


@pytest.mark.parametrize("C", [8, 12, 16, 20, 24])
@pytest.mark.parametrize("n,kind,tol", [(1000, "smooth", 3e-12), (74, "smooth", 3e-12), (300, "tophat", 2e-9),
                                        (1300, "smooth", 3e-12)])
def test_fused_reverse_step_on_dirichlet_rows(emu, C: int, n: int, kind: str, tol: float) -> None:
    """Dirichlet rows (scalar.py:418-427) with data that change from stage to stage (t, t + dt, t + dt / 2,
    timestepping.py:314-319): the ghost cells of u, k1, k2 take the data of their stage inside the kernel and carry
    no cotangent; against reverse-mode differentiation of the reference arithmetic with the same boundary
    function, interior cells (the ghost cells of p' are zero, those of the result are not written)"""
    rng = np.random.default_rng(3 * n + C)
    batch = 2
    u = np.stack([_state(n, kind, 10 * n + b) for b in range(batch)])
    p = rng.standard_normal((batch, n + 2 * G))
    p[:, :G] = 0.0
    p[:, n + G :] = 0.0
    dt = 0.3 * (3.0 / n)
    grid = po.make_grid(-1.5, 1.5, n, G)
    scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53", EPS))
    xg = np.concatenate([grid.x[:G], grid.x[-G:]])
    amp = rng.uniform(0.2, 0.9, size=batch)

    def data(b: int, t: float) -> np.ndarray:
        return amp[b] * np.cos(2.0 * xg - 40.0 * t) + 0.1 * b

    # block s of ghost3: the data of every row at the stage time s (t, t + dt, t + dt / 2), rows 2 G apart
    ghost3 = np.stack([np.stack([data(b, t) for b in range(batch)]) for t in (0.0, dt, 0.5 * dt)]).copy()
    out, k1, k2 = _run(emu, C, n, u, p, dt, stages=True, ghost3=ghost3)
    i = slice(G, G + n)
    assert np.isfinite(out[:, i]).all()
    assert np.isnan(out[:, :G]).all() and np.isnan(out[:, n + G :]).all()
    for b in range(batch):
        bc = po.Dirichlet(lambda t, x, b=b: amp[b] * np.cos(2.0 * x - 40.0 * t) + 0.1 * b)
        ref = tt.step_vjp(scheme, grid, bc, dt, 0.0, u[b], p[b])
        err = np.abs(out[b, i] - ref[i]).max() / np.abs(ref[i]).max()
        assert err < tol, (b, err)
        # the recomputed stage values next to the row ends saw the boundary data of their stage
        w = po.apply_boundary(bc, grid, 0.0, u[b])
        r1 = u[b] + dt * po.apply_operator(scheme, grid, bc, 0.0, w)
        assert np.abs(k1[b, i] - r1[i]).max() < 1e-13 * np.abs(r1).max()


def test_fused_reverse_step_on_dirichlet_rows_with_time_independent_data(emu) -> None:
    """ghost_block = 0: one set of boundary data serves the three stage times (psk_ssprk33_step_adjoint on Dirichlet
    rows with the descriptor's data, psk_ssprk33_step_adjoint_bc with ghost3 = NULL)"""
    n, C, batch = 300, 12, 2
    rng = np.random.default_rng(11)
    u = np.stack([_state(n, "smooth", 3 * n + b) for b in range(batch)])
    p = rng.standard_normal((batch, n + 2 * G))
    p[:, :G] = 0.0
    p[:, n + G :] = 0.0
    dt = 0.3 * (3.0 / n)
    ghost = rng.uniform(-0.4, 0.4, size=(batch, 2 * G))
    nx = n + 2 * G
    bufs = []
    for src in (u, p, None):
        a, col0, ld = _aligned(batch, n)
        v = a[:, col0 : col0 + nx]
        v[:] = np.nan if src is None else src
        bufs.append(v)
    U, P, OUT = bufs
    dts = np.full(batch, dt)
    assert emu.emu_reverse_step(C, n, G, batch, ld, 3.0 / n, EPS, _p(U), _p(P), _p(dts), 1, _p(OUT), None, None, 0,
                                _p(ghost), 2 * G, 0) == 0
    grid = po.make_grid(-1.5, 1.5, n, G)
    scheme = po.Scheme("burgers", "rusanov", po.make_reconstruction("wenojs53", EPS))
    xg = np.concatenate([grid.x[:G], grid.x[-G:]])
    i = slice(G, G + n)
    for b in range(batch):
        bc = po.Dirichlet(ga=lambda t, x, b=b: np.interp(x, xg, ghost[b]))
        ref = tt.step_vjp(scheme, grid, bc, dt, 0.0, u[b], p[b])
        assert np.abs(OUT[b, i] - ref[i]).max() / np.abs(ref[i]).max() < 3e-12
