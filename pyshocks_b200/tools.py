"""Small host-side helpers the reference keeps in ``pyshocks/tools.py`` and that the
convergence tests of the hot path use: the EOC recorder (:184-262) and the least-squares
order estimate (:265-290)."""

from __future__ import annotations

from typing import Any

import numpy as np


def estimate_order_of_convergence(x: np.ndarray, y: np.ndarray) -> tuple[float, float]:
    """Least-squares fit of ``y ~ c x^p``; returns ``(c, p)``."""
    if x.size <= 1:
        raise RuntimeError("Need at least two values to estimate order.")
    eps = np.finfo(x.dtype).eps
    c = np.polyfit(np.log10(x + eps), np.log10(y + eps), 1)
    return float(10 ** c[-1]), float(c[-2])


class EOCRecorder:
    def __init__(self, *, name: str = "Error") -> None:
        self.name = name
        self.history: list[tuple[float, float]] = []

    def add_data_point(self, h: Any, error: Any) -> None:
        self.history.append((float(h), float(error)))

    @property
    def estimated_order(self) -> float:
        if not self.history:
            return float("nan")
        h, error = np.array(self.history, dtype=np.float64).T
        return estimate_order_of_convergence(h, error)[1]

    @property
    def max_error(self) -> float:
        return max((e for _, e in self.history), default=0.0)

    def satisfied(self, order: float, atol: float | None = None, *, slack: float = 0) -> bool:
        if not self.history:
            return True
        error = np.array([e for _, e in self.history])
        if atol is None:
            atol = 1.0e2 * float(np.finfo(np.float64).eps)
        return bool(self.estimated_order >= (order - slack) or error.max() < atol)

    def __str__(self) -> str:
        lines = [f"{'h':>12s} {self.name:>14s}"] + [f"{h:12.5e} {e:14.6e}" for h, e in self.history]
        return "\n".join(lines) + f"\n estimated order {self.estimated_order:.3f}"
