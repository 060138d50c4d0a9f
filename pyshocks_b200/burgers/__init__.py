"""Burgers' equation (``pyshocks/burgers/__init__.py:62-98``)."""

from __future__ import annotations

from dataclasses import fields
from typing import Any

from .schemes import (
    ESWENO32,
    BurgersScheme,
    EngquistOsher,
    FiniteVolumeScheme,
    Godunov,
    LaxFriedrichs,
    Rusanov,
)

_SCHEMES: dict[str, type[BurgersScheme]] = {
    "default": LaxFriedrichs,
    "godunov": Godunov,
    "rusanov": Rusanov,
    "lf": LaxFriedrichs,
    "eo": EngquistOsher,
    "esweno32": ESWENO32,
}


def scheme_ids() -> tuple[str, ...]:
    return tuple(_SCHEMES.keys())


def make_scheme_from_name(name: str, **kwargs: Any) -> BurgersScheme:
    cls = _SCHEMES.get(name)
    if cls is None:
        raise ValueError(f"Scheme {name!r} not found. Try one of {', '.join(scheme_ids())}.")
    return cls(**{f.name: kwargs[f.name] for f in fields(cls) if f.name in kwargs and f.init})


__all__ = (
    "ESWENO32", "BurgersScheme", "EngquistOsher", "FiniteVolumeScheme", "Godunov", "LaxFriedrichs", "Rusanov",
    "make_scheme_from_name", "scheme_ids",
)
