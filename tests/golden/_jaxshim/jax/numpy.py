"""``jax.numpy`` facade over NumPy (see the package docstring: fixture tooling only)."""

from __future__ import annotations

import functools
from typing import Any

import numpy as _np

pi = _np.pi
inf = _np.inf
s_ = _np.s_
float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
dtype = _np.dtype
finfo = _np.finfo
ndarray = _np.ndarray
linalg = _np.linalg


def _re(x: Any) -> Any:
    x = _np.asarray(x)
    return x.real if _np.iscomplexobj(x) else x


class _AtIndexer:
    def __init__(self, ary: "ShimArray", idx: Any = None) -> None:
        self.ary = ary
        self.idx = idx

    def __getitem__(self, idx: Any) -> "_AtIndexer":
        return _AtIndexer(self.ary, idx)

    def _copy_for(self, values: Any) -> _np.ndarray:
        out = _np.array(self.ary, copy=True).view(_np.ndarray)
        if _np.iscomplexobj(values) and not _np.iscomplexobj(out):
            out = out.astype(_np.complex128)
        return out

    def set(self, values: Any, **kwargs: Any) -> "ShimArray":
        out = self._copy_for(values)
        out[self.idx] = _np.asarray(values)
        return _wrap(out)

    def add(self, values: Any, **kwargs: Any) -> "ShimArray":
        out = self._copy_for(values)
        out[self.idx] += _np.asarray(values)
        return _wrap(out)


class ShimArray(_np.ndarray):
    """ndarray with ``.at[...]`` updates and complex-step aware comparisons."""

    __array_priority__ = 100.0

    @property
    def at(self) -> _AtIndexer:
        return _AtIndexer(self)

    # comparisons order by the real part so that the reference's branches
    # (where(aavg > 0, ...), x < x0, ...) work under complex-step jacfwd
    def __gt__(self, other: Any) -> Any:
        return _np.greater(_re(self), _re(other)).view(_np.ndarray)

    def __lt__(self, other: Any) -> Any:
        return _np.less(_re(self), _re(other)).view(_np.ndarray)

    def __ge__(self, other: Any) -> Any:
        return _np.greater_equal(_re(self), _re(other)).view(_np.ndarray)

    def __le__(self, other: Any) -> Any:
        return _np.less_equal(_re(self), _re(other)).view(_np.ndarray)

    # jax arrays are immutable: ``t += dt`` must rebind, never mutate in place
    # (timestepping.py:150 relies on it: the checkpoint dict keeps the old ``t``)
    def __iadd__(self, other: Any) -> Any:
        return self + other

    def __isub__(self, other: Any) -> Any:
        return self - other

    def __imul__(self, other: Any) -> Any:
        return self * other

    def __itruediv__(self, other: Any) -> Any:
        return self / other

    def __setitem__(self, key: Any, value: Any) -> None:
        raise TypeError("jax arrays are immutable; use .at[...].set(...)")

    def __bool__(self) -> bool:
        return bool(_np.asarray(self).view(_np.ndarray).item())

    def __format__(self, spec: str) -> str:
        if self.ndim == 0:
            return format(_re(self).item(), spec)
        return super().__format__(spec)

    def astype(self, dt: Any, *args: Any, **kwargs: Any) -> Any:  # type: ignore[override]
        base = _np.asarray(self).view(_np.ndarray)
        if _np.iscomplexobj(base) and not _np.issubdtype(_np.dtype(dt), _np.complexfloating):
            base = base.real
        return _wrap(base.astype(dt, *args, **kwargs))


def _wrap(x: Any) -> Any:
    if isinstance(x, ShimArray):
        return x
    if isinstance(x, _np.ndarray):
        return x.view(ShimArray)
    if isinstance(x, _np.generic):
        return _np.asarray(x).view(ShimArray)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _unwrap(x: Any) -> Any:
    if isinstance(x, ShimArray):
        return x.view(_np.ndarray)
    if isinstance(x, (tuple, list)):
        return type(x)(_unwrap(v) for v in x)
    return x


def _lift(fn: Any) -> Any:
    @functools.wraps(fn)
    def wrapper(*args: Any, **kwargs: Any) -> Any:
        return _wrap(fn(*_unwrap(args), **{k: _unwrap(v) for k, v in kwargs.items()}))

    return wrapper


# {{{ functions that need care (complex-step semantics, dtype handling)


def array(obj: Any, dtype: Any = None, **kwargs: Any) -> ShimArray:
    a = _np.asarray(_unwrap(obj))
    if dtype is not None:
        if _np.iscomplexobj(a) and not _np.issubdtype(_np.dtype(dtype), _np.complexfloating):
            # keep the complex-step perturbation alive through jnp.array(x, dtype=f64)
            return _wrap(_np.array(a, dtype=_np.complex128))
        a = a.astype(dtype)
    return _wrap(_np.array(a))


asarray = array


def abs(x: Any) -> ShimArray:  # noqa: A001
    x = _np.asarray(_unwrap(x))
    if _np.iscomplexobj(x):
        return _wrap(x * _np.sign(x.real))
    return _wrap(_np.abs(x))


def sign(x: Any) -> ShimArray:
    return _wrap(_np.sign(_re(_unwrap(x))))


def _select_pair(x: Any, y: Any, take_x: Any, tie: Any) -> ShimArray:
    x, y = _np.broadcast_arrays(_np.asarray(_unwrap(x)), _np.asarray(_unwrap(y)))
    out = _np.where(take_x, x, y)
    if _np.iscomplexobj(out):
        # JAX convention: the gradient of maximum/minimum at ties is split 1/2-1/2
        out = _np.where(tie, 0.5 * (x + y), out)
    return _wrap(out)


def maximum(x: Any, y: Any) -> ShimArray:
    xr, yr = _re(_unwrap(x)), _re(_unwrap(y))
    return _select_pair(x, y, xr > yr, xr == yr)


def minimum(x: Any, y: Any) -> ShimArray:
    xr, yr = _re(_unwrap(x)), _re(_unwrap(y))
    return _select_pair(x, y, xr < yr, xr == yr)


def _extreme(x: Any, pick: Any, axis: Any = None, **kwargs: Any) -> ShimArray:
    x = _np.asarray(_unwrap(x))
    if not _np.iscomplexobj(x):
        return _wrap(pick(x, axis=axis, **kwargs))
    if axis is not None:
        raise NotImplementedError("complex-step max/min along an axis")
    xr = x.real
    m = pick(xr)
    # JAX convention: gradient shared equally between all positions at the extreme
    return _wrap(_np.asarray(_np.mean(x[xr == m])))


def max(x: Any, axis: Any = None, **kwargs: Any) -> ShimArray:  # noqa: A001
    return _extreme(x, _np.max, axis=axis, **kwargs)


def min(x: Any, axis: Any = None, **kwargs: Any) -> ShimArray:  # noqa: A001
    return _extreme(x, _np.min, axis=axis, **kwargs)


amax = max
amin = min


def isfinite(x: Any) -> Any:
    return _np.isfinite(_re(_unwrap(x)))


def full_like(a: Any, fill_value: Any, dtype: Any = None) -> ShimArray:
    a = _np.asarray(_unwrap(a))
    fill_value = _unwrap(fill_value)
    if dtype is None and _np.iscomplexobj(fill_value):
        dtype = _np.complex128
    return _wrap(_np.full_like(a, fill_value, dtype=dtype))


def where(cond: Any, x: Any = None, y: Any = None) -> Any:
    cond = _np.asarray(_unwrap(cond))
    if x is None:
        return _wrap(_np.where(cond))
    return _wrap(_np.where(cond, _unwrap(x), _unwrap(y)))


# }}}


def __getattr__(name: str) -> Any:
    fn = getattr(_np, name)
    if callable(fn) and not isinstance(fn, type):
        return _lift(fn)
    return fn
