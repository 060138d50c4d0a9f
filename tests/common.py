"""Helpers shared by the test modules: build oracle objects from golden cases."""

from __future__ import annotations

import pathlib

import numpy as np

import cases as C
from oracle import pyshocks_oracle as po

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


def load_golden(name: str) -> dict[str, np.ndarray]:
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def oracle_setup(case: C.Case):
    """(scheme, grid, bc) for the NumPy oracle, mirroring make_golden.build()."""
    rec = po.make_reconstruction(case.rec)
    grid = po.make_grid(case.a, case.b, case.n, case.g)
    velocity = None if case.equation == "burgers" else C.velocity_for(case)
    scheme = po.bind(po.Scheme(case.equation, case.flux, rec, alpha=case.alpha, velocity=velocity), grid)
    if case.bc == "periodic":
        bc = po.Periodic()
    else:
        bc = po.Dirichlet(ga=lambda t, x: C.dirichlet_values(case, float(t), x))
    return scheme, grid, bc


def max_rel(a: np.ndarray, b: np.ndarray) -> float:
    scale = max(float(np.max(np.abs(b))), 1.0e-300)
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) / scale
