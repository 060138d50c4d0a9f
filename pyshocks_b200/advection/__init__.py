"""Linear advection equation (``pyshocks/advection/__init__.py:60-98``)."""

from __future__ import annotations

from dataclasses import fields
from typing import Any

from .schemes import AdvectionScheme, FiniteVolumeScheme, Godunov, Upwind, upwind_flux

_SCHEMES: dict[str, type[AdvectionScheme]] = {"default": Godunov, "godunov": Godunov, "upwind": Godunov}


def scheme_ids() -> tuple[str, ...]:
    return tuple(_SCHEMES.keys())


def make_scheme_from_name(name: str, **kwargs: Any) -> AdvectionScheme:
    cls = _SCHEMES.get(name)
    if cls is None:
        raise ValueError(f"Scheme {name!r} not found. Try one of {', '.join(scheme_ids())}.")
    if "velocity" not in kwargs:
        kwargs["velocity"] = None
    return cls(**{f.name: kwargs[f.name] for f in fields(cls) if f.name in kwargs})


__all__ = ("AdvectionScheme", "FiniteVolumeScheme", "Godunov", "Upwind", "make_scheme_from_name", "scheme_ids", "upwind_flux")
