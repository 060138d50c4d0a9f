// psk_common.cuh -- shared host/device plumbing of libpsk (descriptor checks, BC-mapped
// loads, reductions).  See include/psk.h for the ABI contract.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/psk.h"

namespace psk {

constexpr int kSMs = 148;  // B200

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return PSK_E_CUDA;
}

#define PSK_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t e__ = (expr);                               \
    if (e__ != cudaSuccess) return ::psk::cuda_fail(e__);   \
  } while (0)

inline int rec_halo(int rec) {
  return rec == PSK_REC_WENOJS53 ? 3 : ((rec == PSK_REC_WENOJS32 || rec == PSK_REC_ESWENO32) ? 2 : 1);
}

// Rows of a batch over grid.y x grid.z (each <= 65535; a kernel's row is blockIdx.y + blockIdx.z * gridDim.y):
// the smallest z that divides the batch.  false: no such split (the caller slices the batch or falls back).
inline bool split_rows(int batch, unsigned &gy, unsigned &gz) {
  if (batch <= 0) return false;
  for (unsigned z = (static_cast<unsigned>(batch) + 65534u) / 65535u; z <= 65535u && z <= static_cast<unsigned>(batch); ++z)
    if (batch % z == 0) {
      gy = static_cast<unsigned>(batch) / z;
      gz = z;
      return true;
    }
  return false;
}

// Validates what every entry point relies on; returns PSK_OK or an error code.
inline int check_desc(const psk_desc *d) {
  if (d == nullptr) return PSK_E_INVALID;
  if (d->n <= 0 || d->batch <= 0 || d->g < 0) return PSK_E_INVALID;
  if (d->ld < static_cast<int64_t>(d->n) + 2 * d->g) return PSK_E_INVALID;
  if (d->equation < PSK_EQ_BURGERS || d->equation > PSK_EQ_CONTINUITY) return PSK_E_UNSUPPORTED;
  if (d->flux < PSK_FLUX_RUSANOV || d->flux > PSK_FLUX_ESWENO) return PSK_E_UNSUPPORTED;
  if (d->rec < PSK_REC_CONSTANT || d->rec > PSK_REC_ESWENO32) return PSK_E_UNSUPPORTED;
  // "ESWENO32 scheme requires the ESWENO32 reconstruction" (burgers/schemes.py:214-216)
  if (d->flux == PSK_FLUX_ESWENO && d->rec != PSK_REC_ESWENO32) return PSK_E_INVALID;
  if (d->bc < PSK_BC_PERIODIC || d->bc > PSK_BC_NONE) return PSK_E_UNSUPPORTED;
  if (d->math != PSK_MATH_FAST && d->math != PSK_MATH_STRICT) return PSK_E_INVALID;
  // advection / continuity only come with the upwind ("godunov") flux
  if (d->equation != PSK_EQ_BURGERS && d->flux != PSK_FLUX_UPWIND) return PSK_E_UNSUPPORTED;
  // assert grid.nghosts >= rec.stencil_width (reconstruction.py:369, :161)
  if (d->g < rec_halo(d->rec)) return PSK_E_INVALID;
  if (d->bc == PSK_BC_PERIODIC && d->n < d->g) return PSK_E_INVALID;
  if (d->bc == PSK_BC_NEUMANN && d->n < d->g) return PSK_E_INVALID;
  if ((d->bc == PSK_BC_DIRICHLET || d->bc == PSK_BC_NEUMANN) && d->ghost == nullptr)
    return PSK_E_INVALID;
  if (d->equation != PSK_EQ_BURGERS &&
      (d->velocity == nullptr || d->vel_l == nullptr || d->vel_r == nullptr))
    return PSK_E_INVALID;
  if (!(d->dx > 0.0)) return PSK_E_INVALID;
  return PSK_OK;
}

// Everything a kernel needs to evaluate w = apply_boundary(u) at an arbitrary cell.
struct BcView {
  const double *ghost;  // row 0 of the ghost data (or nullptr)
  int64_t ghost_ld;
  int bc, n, g, nx;
};

inline BcView make_bc_view(const psk_desc *d) {
  BcView v;
  v.ghost = d->ghost;
  v.ghost_ld = d->ghost_ld;
  v.bc = d->bc;
  v.n = d->n;
  v.g = d->g;
  v.nx = d->n + 2 * d->g;
  return v;
}

// w[i] for array index i of one row: zero beyond the array ends (the zero padding of
// jnp.convolve(..., "same"), convolve.py:113-114), boundary data in the ghost cells
// (scalar.py:418-427, :472-500, :529-540), the stored value in the interior.
__device__ __forceinline__ double load_w(const BcView &b, const double *__restrict__ urow, int row,
                                         int i) {
  if (i < 0 || i >= b.nx) return 0.0;
  const int g = b.g;
  if (i >= g && i < b.nx - g) return urow[i];
  switch (b.bc) {
    case PSK_BC_PERIODIC:
      return urow[i < g ? i + b.n : i - b.n];
    case PSK_BC_DIRICHLET: {
      const double *gh = b.ghost + static_cast<int64_t>(row) * b.ghost_ld;
      return gh[i < g ? i : i - b.n];
    }
    case PSK_BC_NEUMANN: {
      const double *gh = b.ghost + static_cast<int64_t>(row) * b.ghost_ld;
      return i < g ? urow[2 * g - 1 - i] + gh[i]
                   : urow[2 * (b.nx - g) - 1 - i] + gh[i - b.n];
    }
    default:
      return urow[i];
  }
}

// max over a warp / block of non-negative doubles with NaN propagation: compare the bit
// patterns as unsigned integers (order-preserving for x >= 0; NaN patterns sort above +inf).
__device__ __forceinline__ unsigned long long abs_bits(double x) {
  return static_cast<unsigned long long>(__double_as_longlong(fabs(x)));
}

__device__ __forceinline__ unsigned long long warp_max_bits(unsigned long long v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
    v = o > v ? o : v;
  }
  return v;
}

}  // namespace psk
