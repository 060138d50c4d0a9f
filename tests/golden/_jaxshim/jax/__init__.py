"""NumPy-backed stand-in for the tiny part of the ``jax`` API that pyshocks uses.

TEST-FIXTURE TOOLING ONLY.  JAX is not installable in this environment (no
network, no wheel), so the reference (``/root/reference/src/pyshocks``) cannot
be imported as shipped.  This shim lets the reference's *own, unmodified* Python
run on NumPy fp64 so that ``tests/golden/make_golden.py`` can record golden
input/output vectors of the hot path.  It is never imported by the product
(``pyshocks_b200``), by ``bench.py`` or by any test that runs on the GPU box.

What differs from real JAX: arithmetic is NumPy's (IEEE fp64, NumPy's
accumulation order in ``convolve``) instead of XLA-CPU's, ``jit`` is the
identity, and ``jacfwd`` is evaluated by complex-step differentiation (exact to
round-off for the piecewise-analytic code on the path; derivative conventions at
kinks follow JAX: ``abs'(0) = 0``, ``maximum`` ties split 1/2-1/2, ``where``
passes the selected branch only).
"""

from __future__ import annotations

import functools
import warnings

import numpy as np

from . import numpy as jnp  # noqa: F401  (jax.numpy)
from .numpy import ShimArray, _wrap

Array = ShimArray


class _Config:
    def __init__(self) -> None:
        self.values: dict[str, object] = {}

    def update(self, name: str, val: object = None, **kwargs: object) -> None:
        if "val" in kwargs:
            val = kwargs["val"]
        self.values[name] = val


config = _Config()


def jit(fun=None, **kwargs):  # noqa: ANN001, ANN201
    if fun is None:
        return functools.partial(jit, **kwargs)
    return fun


def device_put(x):  # noqa: ANN001, ANN201
    return _wrap(np.asarray(x))


def device_get(x):  # noqa: ANN001, ANN201
    return np.asarray(x)


def jacfwd(fun, argnums: int = 0, *, h: float = 1.0e-40):  # noqa: ANN001, ANN201
    """Forward-mode Jacobian by complex-step differentiation."""

    def jac(*args):  # noqa: ANN002, ANN202
        x = np.asarray(args[argnums], dtype=np.float64)
        cols = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", np.exceptions.ComplexWarning)
            for k in range(x.size):
                xc = x.astype(np.complex128)
                xc[k] += 1j * h
                cargs = list(args)
                cargs[argnums] = _wrap(xc)
                y = np.asarray(fun(*cargs))
                cols.append(np.imag(y) / h)
        return _wrap(np.stack(cols, axis=1))

    return jac


class _Lax:
    @staticmethod
    def fori_loop(lo, hi, body, init):  # noqa: ANN001, ANN205
        val = init
        for i in range(lo, hi):
            val = body(i, val)
        return val


lax = _Lax()
