"""Case table shared by ``make_golden.py`` (which runs the reference) and the
parity tests (which rebuild the same inputs for the oracle / the CUDA path).

Everything here is plain data + NumPy so that it works with or without the
reference being importable.
"""

from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Case:
    equation: str  # burgers | advection | continuity
    flux: str  # rusanov | lf | godunov | eo | esweno32
    rec: str  # constant | wenojs32 | wenojs53 | esweno32
    bc: str  # periodic | dirichlet
    alpha: float = 1.0
    n: int = 64
    a: float = -1.5
    b: float = 1.5
    velocity: str = "one"  # one | varying   (advection / continuity only)
    state: str = "smooth"  # smooth | rough
    t: float = 0.125

    @property
    def g(self) -> int:
        return {"constant": 1, "wenojs32": 2, "wenojs53": 3, "esweno32": 2}[self.rec]

    @property
    def key(self) -> str:
        parts = [self.equation, self.flux, self.rec, self.bc, self.state]
        if self.equation != "burgers":
            parts.append(f"v{self.velocity}")
        if self.alpha != 1.0:
            parts.append(f"alpha{self.alpha:g}")
        return "_".join(parts)


def rhs_cases() -> list[Case]:
    cases: list[Case] = []
    for flux, rec, bc, state in itertools.product(
        ["rusanov", "lf", "godunov", "eo"],
        ["wenojs53", "wenojs32", "constant"],
        ["periodic", "dirichlet"],
        ["smooth", "rough"],
    ):
        cases.append(Case("burgers", flux, rec, bc, state=state))
    for flux, rec in itertools.product(["rusanov", "lf"], ["wenojs53", "constant"]):
        cases.append(Case("burgers", flux, rec, "dirichlet", alpha=0.98, state="rough"))
    for eq, rec, bc, vel, state in itertools.product(
        ["advection", "continuity"],
        ["wenojs53", "wenojs32", "constant"],
        ["periodic", "dirichlet"],
        ["one", "varying"],
        ["smooth", "rough"],
    ):
        cases.append(Case(eq, "godunov", rec, bc, a=-1.0, b=1.0, velocity=vel, state=state))
    # ESWENO32 (appended so that the keys above keep their order): the Burgers scheme with its
    # dissipative flux (burgers/schemes.py:205-256), and the reconstruction alone behind the
    # upwind fluxes (tests/test_convergence.py:402 of the reference runs advection this way)
    for bc, state in itertools.product(["periodic", "dirichlet"], ["smooth", "rough"]):
        cases.append(Case("burgers", "esweno32", "esweno32", bc, state=state))
    for eq, bc, vel, state in itertools.product(
        ["advection", "continuity"], ["periodic", "dirichlet"], ["one", "varying"], ["smooth", "rough"]
    ):
        cases.append(Case(eq, "godunov", "esweno32", bc, a=-1.0, b=1.0, velocity=vel, state=state))
    cases.append(Case("burgers", "godunov", "esweno32", "periodic", state="rough"))
    return cases


def grid_arrays(case: Case) -> tuple[np.ndarray, np.ndarray, float]:
    """(x, f, h) exactly as make_uniform_cell_grid builds them (grid.py:158-161)."""
    h = (case.b - case.a) / case.n
    f = np.linspace(case.a - case.g * h, case.b + case.g * h, case.n + 2 * case.g + 1, dtype=np.float64)
    x = (f[1:] + f[:-1]) / 2
    return x, f, h


def state_for(case: Case) -> np.ndarray:
    """Deterministic input state of length nx (ghost entries hold junk on purpose:
    apply_boundary must overwrite them)."""
    x, _, _ = grid_arrays(case)
    xh = (x - case.a) / (case.b - case.a)
    rng = np.random.default_rng(abs(hash_key(case.key)) % (2**32))
    if case.state == "smooth":
        u = 0.3 + np.sin(2.0 * np.pi * xh) + 0.25 * np.cos(6.0 * np.pi * xh + 0.4)
    else:
        u = np.where((xh > 0.3) & (xh < 0.62), 1.0, -0.4) + 0.05 * rng.standard_normal(x.size)
        u[(xh > 0.75) & (xh < 0.85)] = 0.0  # exact zeros: sonic / tie handling
    u = u.copy()
    u[: case.g] = 7.0 + rng.standard_normal(case.g)
    u[x.size - case.g :] = -5.0 + rng.standard_normal(case.g)
    return u


def velocity_for(case: Case) -> np.ndarray:
    x, _, _ = grid_arrays(case)
    if case.velocity == "one":
        return np.ones_like(x)
    xh = (x - case.a) / (case.b - case.a)
    return 0.3 + np.sin(2.0 * np.pi * xh)  # changes sign


def hash_key(key: str) -> int:
    # stable across processes (python's hash() is salted)
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) % (2**64)
    return h


def dirichlet_values(case: Case, t: float, xg: np.ndarray) -> np.ndarray:
    """Ghost-cell Dirichlet data g(t, x) used by the dirichlet cases."""
    return 0.2 + 0.1 * np.sin(3.0 * xg - 2.0 * t)


STEPPERS = ("ForwardEuler", "RK44", "CKRK45")


def stepper_cases() -> list[Case]:
    """cases of ``steppers.npz``: one ``advance`` of every stepper besides SSPRK33; the Dirichlet cases
    exercise the stage times ``t + c_i dt`` (the boundary data depend on time)"""
    return [
        Case("burgers", "rusanov", "wenojs53", "periodic", state="rough"),
        Case("burgers", "lf", "wenojs53", "dirichlet", state="rough"),
        Case("burgers", "eo", "wenojs32", "dirichlet", state="smooth"),
        Case("advection", "godunov", "wenojs53", "dirichlet", velocity="varying", state="rough"),
        Case("continuity", "godunov", "wenojs53", "periodic", velocity="varying", state="smooth"),
    ]
