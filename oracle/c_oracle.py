"""ctypes front-end of ``oracle/libpsk_oracle.so`` (the plain-C restatement).

TEST INFRASTRUCTURE ONLY -- see ``psk_oracle.c``.  All arrays are host NumPy
fp64; the descriptor mirrors ``psk_desc`` of ``include/psk.h``.
"""

from __future__ import annotations

import ctypes as ct
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
LIB = HERE / "libpsk_oracle.so"

EQUATION = {"burgers": 0, "advection": 1, "continuity": 2}
FLUX = {"rusanov": 0, "lf": 1, "godunov": 2, "eo": 3, "esweno32": 4}
REC = {"constant": 0, "wenojs32": 1, "wenojs53": 2, "esweno32": 3}
BC = {"periodic": 0, "dirichlet": 1, "neumann": 2, "none": 3}

_dp = ct.POINTER(ct.c_double)


class Desc(ct.Structure):
    _fields_ = [
        ("equation", ct.c_int32),
        ("flux", ct.c_int32),
        ("rec", ct.c_int32),
        ("bc", ct.c_int32),
        ("math", ct.c_int32),
        ("n", ct.c_int32),
        ("g", ct.c_int32),
        ("batch", ct.c_int32),
        ("ld", ct.c_int64),
        ("dx", ct.c_double),
        ("eps", ct.c_double),
        ("nu", _dp),
        ("velocity", _dp),
        ("vel_l", _dp),
        ("vel_r", _dp),
        ("ghost", _dp),
        ("ghost_ld", ct.c_int64),
        ("delta", ct.c_double),
    ]


def build(force: bool = False) -> pathlib.Path:
    src = HERE / "psk_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "libpsk_oracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib() -> ct.CDLL:
    global _lib
    if _lib is None:
        _lib = ct.CDLL(str(build()))
        _lib.pso_solve_fixed_dt.argtypes = [ct.POINTER(Desc), _dp, ct.c_double, ct.c_int]
        _lib.pso_solve_adaptive.argtypes = [
            ct.POINTER(Desc), _dp, ct.c_double, ct.c_double, ct.c_double, ct.c_int, _dp,
        ]
        _lib.pso_set_threads.argtypes = [ct.c_int]
        _lib.pso_set_threads.restype = ct.c_int
    return _lib


def set_threads(nthreads: int) -> int:
    """OpenMP threads of the row loops (0: leave as is); returns the count in use."""
    return int(lib().pso_set_threads(int(nthreads)))


def _p(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


class COracle:
    """One bound (scheme, grid, bc-kind) on host arrays of shape (batch, nx)."""

    def __init__(
        self,
        *,
        equation: str,
        flux: str,
        rec: str,
        bc: str,
        n: int,
        g: int,
        batch: int,
        dx: float,
        eps: float,
        delta: float = 0.0,
        nu: np.ndarray | None = None,
        velocity: np.ndarray | None = None,
    ) -> None:
        self.nx = n + 2 * g
        self.batch = batch
        self.keep = {"nu": None if nu is None else np.ascontiguousarray(nu, dtype=np.float64)}
        d = Desc()
        d.equation, d.flux, d.rec, d.bc = EQUATION[equation], FLUX[flux], REC[rec], BC[bc]
        d.math = 1
        d.n, d.g, d.batch, d.ld = n, g, batch, self.nx
        d.dx, d.eps, d.delta = dx, eps, delta
        d.nu = _p(self.keep["nu"])
        self.d = d
        if velocity is not None:
            v = np.ascontiguousarray(velocity, dtype=np.float64)
            d1 = Desc()
            ct.memmove(ct.byref(d1), ct.byref(d), ct.sizeof(Desc))
            d1.batch = 1
            vl, vr = np.empty_like(v), np.empty_like(v)
            lib().pso_reconstruct(ct.byref(d1), _p(v), _p(vl), _p(vr))
            self.keep.update(v=v, vl=vl, vr=vr)
            d.velocity, d.vel_l, d.vel_r = _p(v), _p(vl), _p(vr)

    def set_ghost(self, ghost: np.ndarray | None) -> None:
        """ghost: (2g,) shared or (batch, 2g) per row."""
        if ghost is None:
            self.d.ghost, self.d.ghost_ld = None, 0
            return
        gh = np.ascontiguousarray(ghost, dtype=np.float64)
        self.keep["ghost"] = gh
        self.d.ghost = _p(gh)
        self.d.ghost_ld = 0 if gh.ndim == 1 else gh.shape[1]

    def _rows(self, u: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(self.batch, self.nx)
        return u

    def apply_boundary(self, u: np.ndarray) -> np.ndarray:
        u = self._rows(u)
        w = np.empty_like(u)
        lib().pso_apply_boundary(ct.byref(self.d), _p(u), _p(w))
        return w

    def reconstruct(self, f: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        f = self._rows(f)
        fl, fr = np.empty_like(f), np.empty_like(f)
        lib().pso_reconstruct(ct.byref(self.d), _p(f), _p(fl), _p(fr))
        return fl, fr

    def numerical_flux(self, w: np.ndarray) -> np.ndarray:
        w = self._rows(w)
        F = np.empty((self.batch, self.nx + 1))
        lib().pso_numerical_flux(ct.byref(self.d), _p(w), _p(F), ct.c_int64(self.nx + 1))
        return F

    def apply_operator(self, u: np.ndarray) -> np.ndarray:
        u = self._rows(u)
        L = np.empty_like(u)
        lib().pso_apply_operator(ct.byref(self.d), _p(u), _p(L))
        return L

    def max_abs(self, u: np.ndarray, interior_only: bool = True) -> np.ndarray:
        u = self._rows(u)
        out = np.empty(self.batch)
        lib().pso_max_abs(ct.byref(self.d), _p(u), ct.c_int(int(interior_only)), _p(out))
        return out

    def ssprk33_step(self, u: np.ndarray, dt: np.ndarray | float, ghost3: np.ndarray | None = None) -> np.ndarray:
        u = self._rows(u)
        dt = np.atleast_1d(np.asarray(dt, dtype=np.float64))
        out = np.empty_like(u)
        g3 = None if ghost3 is None else np.ascontiguousarray(ghost3, dtype=np.float64)
        lib().pso_ssprk33_step(
            ct.byref(self.d), _p(u), _p(dt), ct.c_int64(0 if dt.size == 1 else 1), _p(g3), _p(out)
        )
        return out

    def solve_fixed_dt(self, u: np.ndarray, dt: float, nsteps: int) -> np.ndarray:
        u = self._rows(u).copy()
        lib().pso_solve_fixed_dt(ct.byref(self.d), _p(u), float(dt), int(nsteps))
        return u

    def solve_adaptive(self, u: np.ndarray, theta: float, cfl_scale: float, tfinal: float, max_steps: int = 1 << 20):
        assert self.batch == 1
        u = self._rows(u).copy()
        hist = np.zeros(max_steps if max_steps < (1 << 16) else (1 << 16))
        m = lib().pso_solve_adaptive(
            ct.byref(self.d), _p(u), float(theta), float(cfl_scale), float(tfinal), int(hist.size), _p(hist)
        )
        if m < 0:
            raise ValueError("Time step is not finite")
        return u[0], hist[:m]
