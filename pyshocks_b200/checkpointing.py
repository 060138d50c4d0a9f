"""Forward-state tape for the adjoint (``pyshocks/checkpointing.py:27-140``).

Only the in-memory checkpoint is on the path: ``step`` saves ``{"m", "t", "u"}`` every step
(timestepping.py:130-131) and ``adjoint_step`` loads them in reverse (:200).  The arrays stay on
the GPU.  The on-disk variants of the reference are outside the hot path (and raise when the
file does not exist, checkpointing.py:170-171)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from functools import singledispatch
from typing import Any, Hashable


@dataclass(frozen=True)
class Checkpoint(ABC):
    basename: str

    @abstractmethod
    def index_to_key(self, i: int) -> Hashable:
        """Key of checkpoint *i*."""

    @abstractmethod
    def __contains__(self, i: int) -> bool:
        pass


@singledispatch
def save(chk: Checkpoint, idx: int, values: dict[str, Any]) -> None:
    raise NotImplementedError(type(chk).__name__)


@singledispatch
def load(chk: Checkpoint, idx: int, *, include: tuple[str, ...] | None = None) -> dict[str, Any]:
    raise NotImplementedError(type(chk).__name__)


@dataclass(frozen=True)
class InMemoryCheckpoint(Checkpoint):
    storage: dict[Hashable, dict[str, Any]] = field(default_factory=dict)

    def index_to_key(self, i: int) -> Hashable:
        return (self.basename, i)

    def __contains__(self, i: int) -> bool:
        return self.index_to_key(i) in self.storage


@save.register(InMemoryCheckpoint)
def save_in_memory(chk: InMemoryCheckpoint, idx: int, values: dict[str, Any]) -> None:
    key = chk.index_to_key(idx)
    if key in chk.storage:
        raise KeyError(f"Cannot set existing checkpoint at {idx!r}.")
    chk.storage[key] = values


@load.register(InMemoryCheckpoint)
def load_in_memory(chk: InMemoryCheckpoint, idx: int, *, include: tuple[str, ...] | None = None) -> dict[str, Any]:
    key = chk.index_to_key(idx)
    if key not in chk.storage:
        raise KeyError(f"Cannot find checkpoint at index {idx!r}.")
    if include is None:
        return chk.storage[key]
    return {k: v for k, v in chk.storage[key].items() if k in include}
