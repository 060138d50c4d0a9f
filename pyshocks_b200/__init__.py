"""pyshocks_b200 -- the WENO-JS + SSPRK33 hot path of alexfikl/pyshocks as hand-written fp64
CUDA for B200 (sm_100a), behind pyshocks' own operator API.

The names re-exported here are the in-scope part of ``pyshocks/__init__.py:8-54``.  Arrays are
``torch.float64`` CUDA tensors; every registered implementation launches kernels of
``libpsk.so`` through the C ABI of ``include/psk.h`` (there is no CPU fallback: importing this
package without the built library raises).
"""

from . import _lib  # noqa: F401  (fails loudly when libpsk.so is missing)
from . import advection, burgers, continuity, funcs, reconstruction, timestepping  # noqa: F401
from .binding import NoBoundary
from .grid import (
    Grid,
    Quadrature,
    UniformGrid,
    cell_average,
    make_leggauss_quadrature,
    make_uniform_cell_grid,
    norm,
    rnorm,
)
from .schemes import (
    Boundary,
    BoundaryType,
    ConservationLawScheme,
    FiniteDifferenceSchemeBase,
    FiniteVolumeSchemeBase,
    SchemeBase,
    SchemeT,
    apply_boundary,
    apply_operator,
    bind,
    evaluate_boundary,
    flux,
    numerical_flux,
    predict_timestep,
)
from .timestepping import bind_operator, jit
from .tools import EOCRecorder, estimate_order_of_convergence

__version__ = "0.1.0"

__all__ = (
    "Boundary", "BoundaryType", "ConservationLawScheme", "EOCRecorder", "FiniteDifferenceSchemeBase",
    "FiniteVolumeSchemeBase", "Grid", "NoBoundary", "Quadrature", "SchemeBase", "SchemeT", "UniformGrid",
    "apply_boundary", "apply_operator", "bind", "bind_operator", "cell_average",
    "estimate_order_of_convergence", "evaluate_boundary", "flux", "jit", "make_leggauss_quadrature",
    "make_uniform_cell_grid", "norm", "numerical_flux", "predict_timestep", "rnorm",
)
