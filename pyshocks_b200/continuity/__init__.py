"""Continuity equation (``pyshocks/continuity/__init__.py:55-84``)."""

from __future__ import annotations

from dataclasses import fields
from typing import Any

from .schemes import ContinuityScheme, FiniteVolumeScheme, Godunov

_SCHEMES: dict[str, type[ContinuityScheme]] = {"default": Godunov, "godunov": Godunov, "upwind": Godunov}


def scheme_ids() -> tuple[str, ...]:
    return tuple(_SCHEMES.keys())


def make_scheme_from_name(name: str, **kwargs: Any) -> ContinuityScheme:
    cls = _SCHEMES.get(name)
    if cls is None:
        raise ValueError(f"Scheme {name!r} not found. Try one of {', '.join(scheme_ids())}.")
    if "velocity" not in kwargs:
        kwargs["velocity"] = None
    return cls(**{f.name: kwargs[f.name] for f in fields(cls) if f.name in kwargs})


__all__ = ("ContinuityScheme", "FiniteVolumeScheme", "Godunov", "make_scheme_from_name", "scheme_ids")
