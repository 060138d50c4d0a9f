"""ctypes binding of ``libpsk.so`` -- the thin C-ABI layer between the Python host
code and the hand-written sm_100a kernels (``include/psk.h``).

There is no CPU fallback: if the shared library is missing this module raises at
import time, and every call raises :class:`PskError` on a non-zero status.
"""

from __future__ import annotations

import ctypes as ct
import os
import pathlib

import torch

CSRC = pathlib.Path(__file__).resolve().parent / "csrc"
LIB_PATH = pathlib.Path(os.environ["PSK_LIB"]) if os.environ.get("PSK_LIB") else CSRC / "libpsk.so"  # PSK_LIB: A/B builds

# enum values of include/psk.h
EQ_BURGERS, EQ_ADVECTION, EQ_CONTINUITY = 0, 1, 2
FLUX_RUSANOV, FLUX_LAX_FRIEDRICHS, FLUX_UPWIND, FLUX_ENGQUIST_OSHER, FLUX_ESWENO = 0, 1, 2, 3, 4
REC_CONSTANT, REC_WENOJS32, REC_WENOJS53, REC_ESWENO32 = 0, 1, 2, 3
BC_PERIODIC, BC_DIRICHLET, BC_NEUMANN, BC_NONE = 0, 1, 2, 3
MATH_FAST, MATH_STRICT = 0, 1

OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_NONFINITE = 0, 1, 2, 3, 4

_dp = ct.c_void_p


class PskDesc(ct.Structure):
    """``psk_desc`` of include/psk.h (device pointers as integers)."""

    _fields_ = [
        ("equation", ct.c_int32),
        ("flux", ct.c_int32),
        ("rec", ct.c_int32),
        ("bc", ct.c_int32),
        ("math", ct.c_int32),
        ("n", ct.c_int32),
        ("g", ct.c_int32),
        ("batch", ct.c_int32),
        ("ld", ct.c_int64),
        ("dx", ct.c_double),
        ("eps", ct.c_double),
        ("nu", _dp),
        ("velocity", _dp),
        ("vel_l", _dp),
        ("vel_r", _dp),
        ("ghost", _dp),
        ("ghost_ld", ct.c_int64),
        ("delta", ct.c_double),
    ]


class PskHaloLink(ct.Structure):
    """``psk_halo_link`` of include/psk.h (device pointers as integers)."""

    _fields_ = [
        ("wait_lo", _dp),
        ("wait_hi", _dp),
        ("wait_epoch", ct.c_int64),
        ("peer_lo", _dp),
        ("peer_hi", _dp),
        ("flag_lo", _dp),
        ("flag_hi", _dp),
        ("timeout_ns", ct.c_int64),
        ("timed_out", _dp),
        ("epoch_in", _dp),
        ("epoch_out", _dp),
    ]


class PskError(RuntimeError):
    def __init__(self, fn: str, status: int) -> None:
        detail = _lib.psk_status_string(status).decode()
        if status == E_CUDA:
            detail += f" (cudaError {_lib.psk_last_cuda_error()})"
        super().__init__(f"{fn} failed: {detail} [status {status}]")
        self.status = status


def _load() -> ct.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). "
            "pyshocks_b200 has no CPU fallback."
        )
    lib = ct.CDLL(str(LIB_PATH))
    D = ct.POINTER(PskDesc)
    i32, i64, f64, vp = ct.c_int32, ct.c_int64, ct.c_double, ct.c_void_p
    sigs = {
        "psk_version": ([], ct.c_int),
        "psk_status_string": ([ct.c_int], ct.c_char_p),
        "psk_last_cuda_error": ([], ct.c_int),
        "psk_set_stage_variant": ([ct.c_int], ct.c_int),
        "psk_set_adjoint_variant": ([ct.c_int], ct.c_int),
        "psk_apply_boundary": ([D, vp, vp, vp], ct.c_int),
        "psk_reconstruct": ([D, vp, vp, vp, vp], ct.c_int),
        "psk_numerical_flux": ([D, vp, vp, i64, vp, vp], ct.c_int),
        "psk_apply_operator": ([D, vp, vp, vp, vp], ct.c_int),
        "psk_max_abs": ([D, vp, ct.c_int, vp, vp], ct.c_int),
        "psk_ssprk33_stage": ([D, ct.c_int, vp, vp, vp, vp, i64, vp, vp, vp, ct.c_int, vp], ct.c_int),
        "psk_rhs_axpby": ([D, vp, vp, vp, vp, i64, f64, f64, f64, vp, ct.c_int, vp], ct.c_int),
        "psk_ssprk33_stage_lf": ([D, ct.c_int, vp, vp, vp, vp, i64, vp, vp, vp], ct.c_int),
        "psk_ssprk33_step": ([D, vp, vp, vp, i64, vp, vp, vp], ct.c_int),
        "psk_ssprk33_step_bc": ([D, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp], ct.c_int),
        "psk_ssprk33_steps_tape": ([D, vp, i64, ct.c_int, vp, vp, vp], ct.c_int),
        "psk_ssprk33_step_stages": ([D, vp, vp, vp, vp, vp, i64, vp], ct.c_int),
        "psk_ssprk33_step_adjoint": ([D, vp, vp, vp, i64, vp, vp, vp, vp], ct.c_int),
        "psk_ssprk33_step_adjoint_bc": ([D, vp, vp, vp, i64, vp, vp, vp, vp, vp], ct.c_int),
        "psk_set_reverse_variant": ([ct.c_int], ct.c_int),
        "psk_step_control": ([i32, f64, f64, f64, vp, vp, vp, vp, vp, vp, vp], ct.c_int),
        "psk_solve_rows": ([D, vp, ct.c_int, f64, f64, f64, f64, ct.c_int, vp, vp, vp, vp, vp], ct.c_int),
        "psk_solve_rows_tables": ([D, vp, ct.c_int, vp, vp, vp, vp, vp, vp], ct.c_int),
        "psk_ssprk33_adjoint_sweep": ([D, vp, i64, ct.c_int, vp, vp, D, vp, vp, vp, vp, vp, vp], ct.c_int),
        "psk_dfma_probe": ([vp, ct.c_int, ct.c_int, vp], ct.c_int),
        "psk_apply_operator_vjp": ([D, vp, vp, vp, vp, vp], ct.c_int),
        "psk_ssprk33_stage_adjoint": (
            [D, vp, vp, vp, i64, f64, vp, f64, vp, f64, vp, vp, vp], ct.c_int
        ),
        "psk_p2p_alloc": ([ct.c_uint64, ct.POINTER(vp), ct.c_char_p], ct.c_int),
        "psk_p2p_free": ([vp], ct.c_int),
        "psk_p2p_open": ([ct.c_char_p, ct.POINTER(vp)], ct.c_int),
        "psk_p2p_close": ([vp], ct.c_int),
        "psk_halo_push": ([vp, vp, vp, vp, i32, vp, vp, i64, vp], ct.c_int),
        "psk_halo_wait": ([vp, vp, i64, i64, vp, vp], ct.c_int),
        "psk_ssprk33_stage_p2p": ([D, ct.c_int, vp, vp, vp, vp, vp, ct.POINTER(PskHaloLink), vp], ct.c_int),
        "psk_ssprk33_step_p2p": ([D, vp, vp, vp, vp, ct.POINTER(PskHaloLink), vp], ct.c_int),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = the library is stale: rebuild it
        fn.argtypes = argtypes
        fn.restype = restype
    return lib


_lib = _load()
if os.environ.get("PSK_ADJOINT_VARIANT"):  # A/B measurements only
    _lib.psk_set_adjoint_variant(int(os.environ["PSK_ADJOINT_VARIANT"]))
if os.environ.get("PSK_REVERSE_VARIANT"):  # A/B measurements only
    _lib.psk_set_reverse_variant(int(os.environ["PSK_REVERSE_VARIANT"]))
if os.environ.get("PSK_STAGE_VARIANT"):  # A/B measurements only
    _lib.psk_set_stage_variant(int(os.environ["PSK_STAGE_VARIANT"]))
EXPORTS = (
    "psk_version", "psk_status_string", "psk_last_cuda_error", "psk_set_stage_variant", "psk_set_adjoint_variant", "psk_apply_boundary",
    "psk_reconstruct", "psk_numerical_flux", "psk_apply_operator", "psk_max_abs",
    "psk_ssprk33_stage", "psk_rhs_axpby", "psk_ssprk33_stage_lf", "psk_ssprk33_step", "psk_ssprk33_step_bc", "psk_ssprk33_steps_tape", "psk_ssprk33_step_stages", "psk_ssprk33_step_adjoint", "psk_ssprk33_step_adjoint_bc", "psk_set_reverse_variant", "psk_step_control", "psk_solve_rows", "psk_solve_rows_tables", "psk_ssprk33_adjoint_sweep", "psk_dfma_probe", "psk_apply_operator_vjp",
    "psk_ssprk33_stage_adjoint", "psk_p2p_alloc", "psk_p2p_free", "psk_p2p_open", "psk_p2p_close",
    "psk_halo_push", "psk_halo_wait", "psk_ssprk33_stage_p2p", "psk_ssprk33_step_p2p",
)


def lib() -> ct.CDLL:
    return _lib


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a CUDA fp64 tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TypeError("pyshocks_b200 arrays live on the GPU (torch CUDA tensors); got a CPU tensor")
    if t.dtype != torch.float64:
        raise TypeError(f"expected float64, got {t.dtype}")
    return t.data_ptr()


def raw_ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of any CUDA tensor (masks, flags)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TypeError("expected a CUDA tensor")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def check(fn: str, status: int) -> None:
    if status != OK:
        raise PskError(fn, status)


def rows_of(u: torch.Tensor) -> tuple[int, int, int]:
    """(batch, nx, ld) of a state array: 1-D (nx,) or 2-D (batch, nx) with unit inner stride."""
    if u.dim() == 1:
        if u.stride(0) != 1:
            raise ValueError("state arrays must have unit stride along x")
        return 1, u.shape[0], u.shape[0]
    if u.dim() == 2:
        if u.stride(1) != 1 and u.shape[1] != 1:
            raise ValueError("state arrays must have unit stride along x")
        ld = u.stride(0) if u.shape[0] > 1 else max(u.stride(0), u.shape[1])
        return u.shape[0], u.shape[1], ld
    raise ValueError(f"state arrays are 1-D (nx,) or 2-D (batch, nx); got shape {tuple(u.shape)}")
