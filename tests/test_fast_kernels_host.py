"""The specialised stage kernels of the product (pyshocks_b200/csrc/psk_fast_kernels.cuh: the
first 120-cell warp layout, the shared-difference layout with early and late u0 loads, the
whole-step kernel) compiled for the HOST and run under a 32-thread warp emulation
(tests/host/fast_kernels_host.cpp, tests/host/emu/cuda_runtime.h), against the C oracle
(oracle/psk_oracle.c, the restatement of schemes.py:339-346 / scalar.py / reconstruction.py /
timestepping.py:312-320).  What this pins without a GPU: the lane -> cell maps, the halo shuffles
and halo loads, the row tails, the boundary conditions and that every layout performs the SAME
arithmetic per cell wherever the cell sits (bitwise equal outputs, bitwise shift equivariance)."""

from __future__ import annotations

import ctypes as ct
import pathlib
import shutil
import subprocess

import numpy as np
import pytest

from oracle.c_oracle import BC, EQUATION, FLUX, COracle

ROOT = pathlib.Path(__file__).resolve().parent.parent
G = 3
EPS = 1.0e-12


@pytest.fixture(scope="module")
def emu(tmp_path_factory: pytest.TempPathFactory) -> ct.CDLL:
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("emu") / "libfastemu.so"
    subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-pthread",
                    "-I", str(ROOT / "tests" / "host" / "emu"), "-o", str(out),
                    str(ROOT / "tests" / "host" / "fast_kernels_host.cpp")], check=True)
    lib = ct.CDLL(str(out))
    dp, up = ct.POINTER(ct.c_double), ct.POINTER(ct.c_ulonglong)
    lib.emu_fast_stage.argtypes = [ct.c_int] * 10 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, dp,
                                                     ct.c_int, dp, ct.c_longlong, dp, up, dp, dp, dp, ct.c_int]
    lib.emu_fast_stage.restype = ct.c_int
    lib.emu_fused_step.argtypes = [ct.c_int] * 7 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, ct.c_int,
                                                   ct.POINTER(ct.c_ubyte), up]
    lib.emu_fused_step.restype = ct.c_int
    lib.emu_fused_step_stages.argtypes = [ct.c_int] * 3 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, dp, dp, ct.c_int]
    lib.emu_fused_step_stages.restype = ct.c_int
    lib.emu_fused_step_bc.argtypes = [ct.c_int] * 7 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, ct.c_int, dp,
                                                      ct.c_longlong, dp, dp, dp, up, dp]
    lib.emu_fused_step_bc.restype = ct.c_int
    lib.emu_fused_step_periodic_eq.argtypes = [ct.c_int] * 4 + [ct.c_longlong, ct.c_double, ct.c_double, dp, dp, dp, ct.c_int,
                                                               dp, dp, dp]
    lib.emu_fused_step_periodic_eq.restype = ct.c_int
    lib.emu_chunks_per_row.argtypes = [ct.c_int]
    lib.emu_chunks_per_row.restype = ct.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ct.POINTER(ct.c_double))


class Problem:
    """one bound scheme on (batch, nx) host rows + the matching C oracle"""

    def __init__(self, equation: str, flux: str, bc: str, n: int, batch: int, seed: int = 0) -> None:
        self.equation, self.flux, self.bc, self.n, self.batch = equation, flux, bc, n, batch
        self.nx = n + 2 * G
        self.dx = 3.0 / n
        rng = np.random.default_rng(seed)
        x = (np.arange(self.nx) - G + 0.5) / n
        self.u = np.stack([
            rng.uniform(-0.5, 0.5) + sum(rng.uniform(0, 1 / k) * np.sin(2 * np.pi * k * x + rng.uniform(0, 6.28))
                                          for k in range(1, 5))
            for _ in range(batch)
        ])
        # a discontinuity in every row: the nonlinear weights leave their smooth-data values
        self.u[:, G + n // 3 : G + n // 2] += 0.7
        velocity = None
        if equation != "burgers":
            velocity = 1.0 + 0.3 * np.sin(2 * np.pi * x + 0.3)
        self.co = COracle(equation=equation, flux=flux, rec="wenojs53", bc=bc, n=n, g=G, batch=batch,
                          dx=self.dx, eps=EPS, velocity=velocity)
        self.ghost = None
        if bc in ("dirichlet", "neumann"):
            self.ghost = rng.uniform(-0.2, 0.2, size=(batch, 2 * G)) * (1.0 if bc == "dirichlet" else self.dx)
            self.co.set_ghost(self.ghost)
        self.dt = (0.3 * self.dx / np.abs(self.u).max(axis=1)) * rng.uniform(0.5, 1.0, size=batch)

    def stage(self, lib, layout: int, late: int, stage: int, uin, u0, with_max: bool = False, wpc_max: int = 8):
        out = np.full_like(uin, np.nan)
        maxabs = np.zeros(self.batch, dtype=np.uint64)
        lf = None
        if self.flux == "lf":
            lf = np.abs(self.co.apply_boundary(uin)).max(axis=1)
        k = self.co.keep
        rc = lib.emu_fast_stage(layout, late, EQUATION[self.equation], FLUX[self.flux], stage, int(with_max),
                                BC[self.bc], self.n, G, self.batch, self.nx, self.dx, EPS, _p(uin), _p(u0), _p(out),
                                _p(self.dt), 1, _p(self.ghost), 0 if self.ghost is None else 2 * G, _p(lf),
                                maxabs.ctypes.data_as(ct.POINTER(ct.c_ulonglong)), _p(k.get("v")), _p(k.get("vl")),
                                _p(k.get("vr")), wpc_max)
        assert rc == 0
        return out, maxabs.view(np.float64)

    def step(self, lib, layout: int, late: int = 0, with_max: bool = False):
        k1, _ = self.stage(lib, layout, late, 1, self.u, self.u)
        k1 = self.fill(k1)
        k2, _ = self.stage(lib, layout, late, 2, k1, self.u)
        k2 = self.fill(k2)
        out, mx = self.stage(lib, layout, late, 3, k2, self.u, with_max=with_max)
        return out, mx

    def fill(self, a):
        """the kernels write interior cells only; the stored ghost cells are never read for periodic /
        Dirichlet / Neumann rows, so any finite filler will do"""
        a = a.copy()
        a[:, :G] = 123.0
        a[:, self.nx - G :] = -321.0
        return a

    @property
    def interior(self):
        return slice(G, G + self.n)


SCHEMES = [("burgers", "rusanov"), ("burgers", "lf"), ("burgers", "godunov"), ("burgers", "eo"),
           ("advection", "godunov"), ("continuity", "godunov")]


@pytest.mark.parametrize("layout", [0, 2, 12])
@pytest.mark.parametrize("equation,flux", SCHEMES)
@pytest.mark.parametrize("bc", ["periodic", "dirichlet"])
def test_step_matches_oracle(emu, layout: int, equation: str, flux: str, bc: str) -> None:
    pb = Problem(equation, flux, bc, n=250, batch=2, seed=7)
    got, _ = pb.step(emu, layout)
    ref = pb.co.ssprk33_step(pb.u, pb.dt)
    i = pb.interior
    assert np.isfinite(got[:, i]).all()
    err = np.abs(got[:, i] - ref[:, i]).max() / np.abs(ref[:, i]).max()
    assert err < 2e-13, err


@pytest.mark.parametrize("n", [3, 4, 5, 119, 120, 121, 125, 126, 127, 128, 129, 240, 252, 253, 379, 504, 1000])
@pytest.mark.parametrize("bc", ["periodic", "neumann"])
def test_layouts_are_bitwise_equal_at_row_tails(emu, n: int, bc: str) -> None:
    """same arithmetic per cell whatever the lane / chunk a cell falls into, for every tail length"""
    pb = Problem("burgers", "rusanov", bc, n=n, batch=2, seed=n)
    ref = pb.co.ssprk33_step(pb.u, pb.dt)
    i = pb.interior
    base, mx0 = pb.step(emu, 0, with_max=True)
    err = np.abs(base[:, i] - ref[:, i]).max() / np.abs(ref[:, i]).max()
    assert err < 2e-13, err
    assert np.array_equal(mx0, np.abs(base[:, i]).max(axis=1))
    for layout in (2, 12):  # shared differences, u0 loaded early / late
        got, mx = pb.step(emu, layout, 0, with_max=True)
        assert np.array_equal(got[:, i], base[:, i]), layout
        assert np.array_equal(mx, mx0)
        # nothing outside the interior is written
        assert np.isnan(got[:, :G]).all() and np.isnan(got[:, G + n :]).all()


@pytest.mark.parametrize("layout", [0, 2])
def test_stage0_is_the_operator(emu, layout: int) -> None:
    pb = Problem("burgers", "rusanov", "periodic", n=300, batch=1, seed=3)
    got, _ = pb.stage(emu, layout, 0, 0, pb.u, pb.u)
    ref = pb.co.apply_operator(pb.u)
    i = pb.interior
    assert np.abs(got[:, i] - ref[:, i]).max() <= 2e-13 * np.abs(ref[:, i]).max()


@pytest.mark.parametrize("layout", [0, 2])
def test_shift_equivariance_is_bitwise(emu, layout: int) -> None:
    """periodic rows: rolling the data by k cells rolls the result by k cells, bit for bit"""
    pb = Problem("burgers", "rusanov", "periodic", n=504, batch=1, seed=5)
    a, _ = pb.step(emu, layout)
    i = pb.interior
    for k in (1, 2, 3, 125, 377):
        pb2 = Problem("burgers", "rusanov", "periodic", n=504, batch=1, seed=5)
        pb2.u[:, i] = np.roll(pb.u[:, i], k, axis=1)
        b, _ = pb2.step(emu, layout)
        assert np.array_equal(np.roll(a[:, i], k, axis=1), b[:, i]), k


def test_geometry(emu) -> None:
    assert emu.emu_chunks_per_row(4096) == 35
    for n in (1, 120, 121, 240, 241):
        assert emu.emu_chunks_per_row(n) == -(-n // 120)


@pytest.mark.parametrize("R,flux", [(4, "rusanov"), (6, "rusanov"), (8, "rusanov"), (10, "rusanov"), (6, "godunov"),
                                    (6, "eo")])
@pytest.mark.parametrize("n", [5, 16, 107, 108, 109, 172, 236, 237, 250, 472, 1000])
def test_fused_step_is_the_three_stages_bit_for_bit(emu, R: int, flux: str, n: int) -> None:
    """psk_ssprk33_step's kernel (temporal blocking, R cells per lane): same bits as three stage
    launches, nothing written outside the interior, fused max |u'|, inactive rows copied through"""
    pb = Problem("burgers", flux, "periodic", n=n, batch=3, seed=100 + n)
    i = pb.interior
    staged, _ = pb.step(emu, 2)
    ref = pb.co.ssprk33_step(pb.u, pb.dt)
    out = np.full_like(pb.u, np.nan)
    maxabs = np.zeros(pb.batch, dtype=np.uint64)
    active = np.array([1, 0, 1], dtype=np.uint8)
    u_in = pb.fill(pb.u)  # the stored ghost cells are never read
    rc = emu.emu_fused_step(R, FLUX[flux], 0, 1, n, G, pb.batch, pb.nx, pb.dx, EPS, _p(u_in), _p(out), _p(pb.dt), 1,
                            active.ctypes.data_as(ct.POINTER(ct.c_ubyte)),
                            maxabs.ctypes.data_as(ct.POINTER(ct.c_ulonglong)))
    assert rc == 0
    assert np.isnan(out[:, :G]).all() and np.isnan(out[:, G + n :]).all()
    for r in (0, 2):
        assert np.array_equal(out[r, i], staged[r, i])
        assert np.abs(out[r, i] - ref[r, i]).max() <= 2e-13 * np.abs(ref[r, i]).max()
        assert maxabs.view(np.float64)[r] == np.abs(out[r, i]).max()
    assert np.array_equal(out[1, i], pb.u[1, i]) and maxabs[1] == 0


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("n", [344, 1000])
def test_fused_step_on_slabs_with_nine_ghost_cells(emu, world: int, n: int) -> None:
    """boundary kind NONE: a slab's window cells beyond its ends are its stored ghost cells (9 per side,
    the neighbours' edge cells); the slabs together give the bits of the undecomposed periodic row"""
    pb = Problem("burgers", "rusanov", "periodic", n=n, batch=1, seed=n)
    whole = np.full_like(pb.u, np.nan)
    mx = np.zeros(1, dtype=np.uint64)
    none_u8 = ct.POINTER(ct.c_ubyte)()
    assert emu.emu_fused_step(6, FLUX["rusanov"], 0, 0, n, G, 1, pb.nx, pb.dx, EPS, _p(pb.u), _p(whole), _p(pb.dt), 1,
                              none_u8, mx.ctypes.data_as(ct.POINTER(ct.c_ulonglong))) == 0
    interior = pb.u[0, G : G + n]
    g9, first, out = 9, 0, []
    for r in range(world):
        nl = n // world + (1 if r < n % world else 0)
        idx = (np.arange(first - g9, first + nl + g9)) % n  # the slab with its neighbours' cells as ghosts
        slab = np.ascontiguousarray(interior[idx][None, :])
        res = np.full_like(slab, np.nan)
        assert emu.emu_fused_step(6, FLUX["rusanov"], 1, 0, nl, g9, 1, nl + 2 * g9, pb.dx, EPS, _p(slab), _p(res),
                                  _p(pb.dt), 1, none_u8, mx.ctypes.data_as(ct.POINTER(ct.c_ulonglong))) == 0
        assert np.isnan(res[0, :g9]).all() and np.isnan(res[0, g9 + nl :]).all()
        out.append(res[0, g9 : g9 + nl])
        first += nl
    assert np.array_equal(np.concatenate(out), whole[0, G : G + n])


def test_fused_step_at_the_bench_row_length(emu) -> None:
    """the row length of BASELINE.json configs[2] (24 windows of 172 cells per row), default shape"""
    test_fused_step_is_the_three_stages_bit_for_bit(emu, 6, "rusanov", 4096)


@pytest.mark.parametrize("n", [16, 172, 173, 500, 1000])
@pytest.mark.parametrize("with_uout", [True, False])
def test_fused_step_with_stored_stages(emu, n: int, with_uout: bool) -> None:
    """psk_ssprk33_step_stages: k1, k2 (and u') of one launch are the bits of the stage launches; without
    uout the third stage is skipped and nothing else is written"""
    pb = Problem("burgers", "rusanov", "periodic", n=n, batch=2, seed=7 * n)
    i = pb.interior
    k1s, _ = pb.stage(emu, 2, 0, 1, pb.u, pb.u)
    k2s, _ = pb.stage(emu, 2, 0, 2, pb.fill(k1s), pb.u)
    outs, _ = pb.stage(emu, 2, 0, 3, pb.fill(k2s), pb.u)
    k1, k2, out = (np.full_like(pb.u, np.nan) for _ in range(3))
    rc = emu.emu_fused_step_stages(n, G, pb.batch, pb.nx, pb.dx, EPS, _p(pb.fill(pb.u)), _p(k1), _p(k2),
                                   _p(out) if with_uout else None, _p(pb.dt), 1)
    assert rc == 0
    assert np.array_equal(k1[:, i], k1s[:, i]) and np.array_equal(k2[:, i], k2s[:, i])
    for a in (k1, k2, out):
        assert np.isnan(a[:, :G]).all() and np.isnan(a[:, G + n :]).all()
    if with_uout:
        assert np.array_equal(out[:, i], outs[:, i])
    else:
        assert np.isnan(out).all()


@pytest.mark.parametrize("equation,flux", [("burgers", "rusanov"), ("burgers", "godunov"), ("burgers", "eo"),
                                           ("advection", "godunov"), ("continuity", "godunov")])
@pytest.mark.parametrize("n", [5, 16, 171, 172, 173, 250, 1000])
@pytest.mark.parametrize("shared_ghosts", [False, True])
@pytest.mark.parametrize("bc", ["dirichlet", "neumann"])
def test_fused_step_with_boundary_data_is_the_three_stages(emu, equation: str, flux: str, n: int, shared_ghosts: bool,
                                                           bc: str) -> None:
    """psk_ssprk33_step_bc's kernel: rows with Dirichlet / Neumann data that change from stage to stage (the values of
    the user's g(t, x) at t, t + dt, t + dt / 2; Neumann ghost cells also mirror the stage values next to the ends,
    scalar.py:472-500), Burgers fluxes and the advection / continuity upwind flux: the bits of three stage launches
    with those data, the C oracle to round-off, nothing written outside the interior"""
    pb = Problem(equation, flux, bc, n=n, batch=3, seed=300 + n)
    rng = np.random.default_rng(n)
    g3 = rng.uniform(-0.3, 0.3, size=(3, 1 if shared_ghosts else pb.batch, 2 * G)) * (1.0 if bc == "dirichlet" else pb.dx)
    i = pb.interior
    cur = pb.u
    for s in (1, 2, 3):  # three stage launches, each with its own boundary data
        pb.ghost = np.ascontiguousarray(np.broadcast_to(g3[s - 1], (pb.batch, 2 * G)))
        cur, _ = pb.stage(emu, 2, 0, s, pb.fill(cur) if s > 1 else cur, pb.u)
    staged = cur
    ref = pb.co.ssprk33_step(pb.u, pb.dt, ghost3=np.ascontiguousarray(np.broadcast_to(g3, (3, pb.batch, 2 * G))))
    out = np.full_like(pb.u, np.nan)
    maxabs = np.zeros(pb.batch, dtype=np.uint64)
    k = pb.co.keep
    g3c = np.ascontiguousarray(g3)
    rc = emu.emu_fused_step_bc(EQUATION[equation], FLUX[flux], 1, int(bc == "neumann"), n, G, pb.batch, pb.nx, pb.dx, EPS, _p(pb.fill(pb.u)),
                               _p(out), _p(pb.dt), 1, _p(g3c), 0 if shared_ghosts else 2 * G, _p(k.get("v")),
                               _p(k.get("vl")), _p(k.get("vr")), maxabs.ctypes.data_as(ct.POINTER(ct.c_ulonglong)), None)
    assert rc == 0
    assert np.isnan(out[:, :G]).all() and np.isnan(out[:, G + n :]).all()
    assert np.array_equal(out[:, i], staged[:, i])
    assert np.abs(out[:, i] - ref[:, i]).max() <= 2e-13 * np.abs(ref[:, i]).max()
    assert np.array_equal(maxabs.view(np.float64), np.abs(out[:, i]).max(axis=1))


@pytest.mark.parametrize("bc", ["dirichlet", "neumann"])
@pytest.mark.parametrize("n", [16, 171, 173, 250, 1000])
def test_fused_step_with_the_viscosity_of_every_face(emu, bc: str, n: int) -> None:
    """Rusanov with alpha != 1 (scalar.py:231-234: nu = df ** (alpha - 1), one value per face of the array, not all
    equal in the last bit) on rows with boundary data: the whole-step kernel multiplies the speed of every face by
    its own nu, against the C oracle bound to the same array (the bit-identity with the stage launches is a GPU
    test: the general stage kernel is not part of the emulated header)"""
    pb = Problem("burgers", "rusanov", bc, n=n, batch=3, seed=500 + n)
    rng = np.random.default_rng(n)
    nu = (pb.dx ** (0.995 - 1.0)) * (1.0 + 1e-3 * rng.standard_normal(pb.nx - 1))
    pb.co = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc=bc, n=n, g=G, batch=pb.batch, dx=pb.dx, eps=EPS,
                    nu=nu)
    pb.co.set_ghost(pb.ghost)  # (per-row boundary data: the layout ghost3 is read with)
    g3 = rng.uniform(-0.3, 0.3, size=(3, pb.batch, 2 * G)) * (1.0 if bc == "dirichlet" else pb.dx)
    pb.dt = pb.dt / nu.max()
    ref = pb.co.ssprk33_step(pb.u, pb.dt, ghost3=np.ascontiguousarray(g3))
    one = COracle(equation="burgers", flux="rusanov", rec="wenojs53", bc=bc, n=n, g=G, batch=pb.batch, dx=pb.dx, eps=EPS)
    one.set_ghost(pb.ghost)
    assert np.abs(one.ssprk33_step(pb.u, pb.dt, ghost3=np.ascontiguousarray(g3)) - ref).max() > 1e-6  # nu matters
    out = np.full_like(pb.u, np.nan)
    maxabs = np.zeros(pb.batch, dtype=np.uint64)
    i = pb.interior
    g3c, uin = np.ascontiguousarray(g3), pb.fill(pb.u)  # (kept alive across the call)
    rc = emu.emu_fused_step_bc(EQUATION["burgers"], FLUX["rusanov"], 0, int(bc == "neumann"), n, G, pb.batch, pb.nx, pb.dx, EPS,
                               _p(uin), _p(out), _p(pb.dt), 1, _p(g3c), 2 * G, None, None, None,
                               maxabs.ctypes.data_as(ct.POINTER(ct.c_ulonglong)), _p(nu))
    assert rc == 0
    assert np.abs(out[:, i] - ref[:, i]).max() <= 2e-13 * np.abs(ref[:, i]).max()


@pytest.mark.parametrize("equation", ["advection", "continuity"])
@pytest.mark.parametrize("n", [16, 171, 173, 250, 1000])
def test_fused_step_for_advection_and_continuity_on_periodic_rows(emu, equation: str, n: int) -> None:
    """psk_ssprk33_step for the linear equations on periodic rows: with a velocity array whose ghost cells are the
    periodic images of its interior (so that its reconstruction is periodic too) the whole-step kernel -- which
    advances the image cells of a window with the velocity data of the cells they are images of -- gives the bits
    of three stage launches"""
    pb = Problem(equation, "godunov", "periodic", n=n, batch=2, seed=40 + n)
    # a sign-changing velocity with exactly periodic ghost cells, and the oracle re-bound to it
    x = (np.arange(n) + 0.5) / n
    vi = 0.2 + np.sin(2 * np.pi * x + 0.3)
    v = np.concatenate([vi[n - G :], vi, vi[:G]])
    pb.co = COracle(equation=equation, flux="godunov", rec="wenojs53", bc="periodic", n=n, g=G, batch=pb.batch, dx=pb.dx,
                    eps=EPS, velocity=v)
    k = pb.co.keep
    assert k["vr"][G - 1] == k["vr"][G + n - 1] and k["vl"][G + n] == k["vl"][G]
    i = pb.interior
    staged, _ = pb.step(emu, 2)
    ref = pb.co.ssprk33_step(pb.u, pb.dt)
    out = np.full_like(pb.u, np.nan)
    rc = emu.emu_fused_step_periodic_eq(EQUATION[equation], n, G, pb.batch, pb.nx, pb.dx, EPS, _p(pb.fill(pb.u)), _p(out),
                                        _p(pb.dt), 1, _p(k["v"]), _p(k["vl"]), _p(k["vr"]))
    assert rc == 0
    assert np.isnan(out[:, :G]).all() and np.isnan(out[:, G + n :]).all()
    assert np.array_equal(out[:, i], staged[:, i])
    assert np.abs(out[:, i] - ref[:, i]).max() <= 2e-13 * np.abs(ref[:, i]).max()
