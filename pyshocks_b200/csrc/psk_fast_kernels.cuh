// psk_fast_kernels.cuh -- the specialised stage kernels of the hot configuration (WENO-JS5, FAST
// math, nu = 1, all rows active, 16-byte aligned rows) and their launch geometry.  Device code
// only (no launches, no CUDA runtime calls), so that tests/host/ can compile this very file for
// the HOST with a 32-thread warp emulation (tests/test_fast_kernels_host.py) and check the
// indexing, the halo traffic and the boundary handling of every layout without a GPU.
#pragma once

#include <type_traits>

#include "psk_common.cuh"
#include "psk_math.cuh"

namespace psk {

constexpr int kHalo = 3;

// FAST-path flux scaled by kFluxScale<EQ, FLUX> (4 F for Rusanov / Lax-Friedrichs, 2 F for
// the other Burgers fluxes, F for advection / continuity): saves the halvings of u^2 / 2.
template <int EQ, int FLUX>
struct FluxScale {
  static constexpr double value =
      (EQ != PSK_EQ_BURGERS) ? 1.0
                             : ((FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? 4.0 : 2.0);
};

// chunks (warps) per row and warps per CTA of the specialised stage kernels (120 emitted cells per
// warp); the CTA size is the divisor-friendly choice in 4..wpc_max that wastes the fewest warps
struct FastGeometry {
  int chunks_per_row, wpc;
};

inline FastGeometry fast_geometry(int n, int wpc_max) {
  FastGeometry geo;
  geo.chunks_per_row = (n + 119) / 120;
  int wpc = 8, best_waste = 1 << 30;
  for (int w = wpc_max; w >= (wpc_max < 4 ? wpc_max : 4); --w) {
    const int waste = ((geo.chunks_per_row + w - 1) / w) * w - geo.chunks_per_row;
    if (waste < best_waste) { best_waste = waste; wpc = w; }
  }
  if (geo.chunks_per_row < wpc) wpc = geo.chunks_per_row;
  geo.wpc = wpc;
  return geo;
}

// ---------------------------------------------------------------------------
// The hot configuration, specialised: WENO-JS5, FAST math, nu = 1, every row active,
// 16-byte aligned rows.  Same algorithm and data movement as stage_warp_kernel; what differs
// is the bookkeeping: the grid is (chunk groups, rows) so no integer division is needed, the
// stage is a template parameter, the Rusanov speed max(|w_j|, |w_j+1|) is an integer max of the
// bit patterns (2 ALU compares + 2 selects instead of an emulated fp64 max), and the candidate
// offsets reuse the smoothness stencils' linear forms (weno53_pair_lean).
struct FastParams {
  const double *uin;
  const double *u0;
  double *uout;
  const double *dt;
  const double *lf_speed;
  unsigned long long *maxabs;
  const double *vel;
  const double *vel_l;
  const double *vel_r;
  BcView bc;
  int64_t ld;
  double coef;   // 1 / (flux scale * dx)
  double eps9;   // eps / 9
  double ca, cb, cc;  // STAGE 4: uout = ca u0 + cb uin + cc dt L(uin)  (psk_rhs_axpby)
  int dt_stride;
  int chunks_per_row;
};

__device__ __forceinline__ double umax_abs(double a, double b) {
  const unsigned long long x = static_cast<unsigned long long>(__double_as_longlong(a)) & 0x7fffffffffffffffull;
  const unsigned long long y = static_cast<unsigned long long>(__double_as_longlong(b)) & 0x7fffffffffffffffull;
  return __longlong_as_double(static_cast<long long>(x > y ? x : y));
}

// max-magnitude of two NON-POSITIVE doubles (sign bit set) as an integer max of the bit patterns
__device__ __forceinline__ double umax_neg(double a, double b) {
  const unsigned long long x = static_cast<unsigned long long>(__double_as_longlong(a));
  const unsigned long long y = static_cast<unsigned long long>(__double_as_longlong(b));
  return __longlong_as_double(static_cast<long long>(x > y ? x : y));
}

#ifndef PSK_FAST_MIN_BLOCKS
#define PSK_FAST_MIN_BLOCKS 4
#endif
// own cells of a lane: stage input and (stages 2, 3) the step's initial state
struct FastIn {
  double v[4], u0[4];
};

template <int STAGE>
__device__ __forceinline__ void fast_load(const FastParams &p, int row, int c0, int lane, bool inside,
                                          FastIn &in) {
  constexpr int R = 4;
  const int g = p.bc.g, n = p.bc.n;
  const int64_t off = static_cast<int64_t>(row) * p.ld + g + c0;
  const bool emit = (lane >= 1) && (lane <= 30);
  if (inside) {
    const double2 q0 = *reinterpret_cast<const double2 *>(p.uin + off);
    const double2 q1 = *reinterpret_cast<const double2 *>(p.uin + off + 2);
    in.v[0] = q0.x; in.v[1] = q0.y; in.v[2] = q1.x; in.v[3] = q1.y;
  } else {
    const double *__restrict__ urow = p.uin + static_cast<int64_t>(row) * p.ld;
#pragma unroll
    for (int r = 0; r < R; ++r) in.v[r] = load_w(p.bc, urow, row, g + c0 + r);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) in.u0[r] = 0.0;
  if (STAGE >= 2 && emit) {
    if (inside) {
      const double2 q0 = *reinterpret_cast<const double2 *>(p.u0 + off);
      const double2 q1 = *reinterpret_cast<const double2 *>(p.u0 + off + 2);
      in.u0[0] = q0.x; in.u0[1] = q0.y; in.u0[2] = q1.x; in.u0[3] = q1.y;
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (c0 + r >= 0 && c0 + r < n) in.u0[r] = p.u0[off + r];
    }
  }
}

template <int EQ, int FLUX, int STAGE, bool WITH_MAX>
__device__ __forceinline__ void fast_compute_store(const FastParams &p, int row, int c0, int lane, bool inside,
                                                   const FastIn &in, double (&out)[4]) {
  constexpr int R = 4;
  constexpr unsigned kFull = 0xffffffffu;
  const int g = p.bc.g, n = p.bc.n;
  const int64_t off = static_cast<int64_t>(row) * p.ld + g + c0;
  const bool emit = (lane >= 1) && (lane <= 30);
  double v[R + 2 * kHalo];
  double u0v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    v[kHalo + r] = in.v[r];
    u0v[r] = in.u0[r];
  }
  v[0] = __shfl_up_sync(kFull, v[4], 1);
  v[1] = __shfl_up_sync(kFull, v[5], 1);
  v[2] = __shfl_up_sync(kFull, v[6], 1);
  v[7] = __shfl_down_sync(kFull, v[3], 1);
  v[8] = __shfl_down_sync(kFull, v[4], 1);
  v[9] = __shfl_down_sync(kFull, v[5], 1);

  double t[R + 5], pq[R + 4];
#pragma unroll
  for (int k = 0; k < R + 5; ++k) t[k] = __dmul_rn(1.0 / 6.0, v[k + 1] - v[k]);
#pragma unroll
  for (int k = 0; k < R + 4; ++k) {
    const double dd = t[k + 1] - t[k];
    pq[k] = fma((13.0 / 3.0) * dd, dd, p.eps9);
  }
  double ul[R], ur[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = r + kHalo;
    const Weno5Pair o = weno53_pair_lean(v[m], t[m - 2], t[m - 1], t[m], t[m + 1], pq[m - 2], pq[m - 1], pq[m]);
    ul[r] = o.ul;
    ur[r] = o.ur;
  }
  const double ur_left = __shfl_up_sync(kFull, ur[R - 1], 1);
  const double ul_right = __shfl_down_sync(kFull, ul[0], 1);

  const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? p.lf_speed[row] : 0.0;
  double F[R + 1];
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const double urj = (f == 0) ? ur_left : ur[f - 1];
    const double ulp = (f == R) ? ul_right : ul[f];
    if (EQ == PSK_EQ_BURGERS) {
      if (FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
        const double a = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? speed : umax_abs(v[f + kHalo - 1], v[f + kHalo]);
        F[f] = fma(-2.0 * a, ulp - urj, fma(urj, urj, ulp * ulp));  // 4 F
      } else if (FLUX == PSK_FLUX_UPWIND) {
        const double w = (urj + ulp) > 0.0 ? urj : ulp;
        F[f] = __dmul_rn(w, w);  // 2 F; never contracted with the flux difference (same bits in every kernel)
      } else {
        const double vp = fmax(urj, 0.0), vm = fmin(ulp, 0.0);
        F[f] = fma(vp, vp, vm * vm);  // 2 F
      }
    } else {
      const int j = g + c0 + f - 1;
      const bool ok = (j >= 0 && j < p.bc.nx - 1);
      const double arj = ok ? p.vel_r[j] : 0.0, alp = ok ? p.vel_l[j + 1] : 0.0;
      const bool pos = (arj + alp) > 0.0;
      F[f] = (EQ == PSK_EQ_ADVECTION) ? (pos ? urj : ulp) : (pos ? __dmul_rn(arj, urj) : __dmul_rn(alp, ulp));
    }
  }

  double coef = p.coef;
  if (STAGE != 0) coef *= p.dt[static_cast<int64_t>(row) * p.dt_stride];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double dF = F[r] - F[r + 1];
    if (EQ == PSK_EQ_ADVECTION) dF *= (c0 + r >= 0 && c0 + r < n) ? p.vel[g + c0 + r] : 0.0;
    if (STAGE == 0) {
      out[r] = coef * dF;
    } else if (STAGE == 4) {
      out[r] = fma(p.cc * coef, dF, fma(p.ca, u0v[r], p.cb * v[r + kHalo]));
    } else {
      const double k = fma(coef, dF, v[r + kHalo]);
      out[r] = (STAGE == 1) ? k
                            : ((STAGE == 2) ? fma(0.25, k, 0.75 * u0v[r]) : fma(2.0 / 3.0, k, (1.0 / 3.0) * u0v[r]));
    }
  }
  if (emit) {
    if (inside) {
      *reinterpret_cast<double2 *>(p.uout + off) = make_double2(out[0], out[1]);
      *reinterpret_cast<double2 *>(p.uout + off + 2) = make_double2(out[2], out[3]);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (c0 + r >= 0 && c0 + r < n) p.uout[off + r] = out[r];
    }
  }
  if (WITH_MAX) {
    unsigned long long mx = 0ull;
    if (emit) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (c0 + r >= 0 && c0 + r < n) {
          const unsigned long long b = abs_bits(out[r]);
          mx = b > mx ? b : mx;
        }
    }
    mx = warp_max_bits(mx);
    if (lane == 0) atomicMax(p.maxabs + row, mx);
  }
}

template <int EQ, int FLUX, int STAGE, bool WITH_MAX>
__global__ void __launch_bounds__(256, PSK_FAST_MIN_BLOCKS)
stage_warp_fast_kernel(const FastParams p) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chunk >= p.chunks_per_row) return;
  const int row = blockIdx.y + blockIdx.z * gridDim.y;
  const int c0 = chunk * 30 * R - R + R * lane;
  const bool inside = (c0 >= 0) && (c0 + R <= p.bc.n);
  FastIn in;
  double out[4];
  fast_load<STAGE>(p, row, c0, lane, inside, in);
  fast_compute_store<EQ, FLUX, STAGE, WITH_MAX>(p, row, c0, lane, inside, in, out);
}

// u0 of a lane's aligned quad (stages 2 and 3); `left` = cells of the quad that exist (n - c0)
template <int STAGE>
__device__ __forceinline__ void fast_load_u0(const double *__restrict__ src, bool inside, int left, double (&u0v)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) u0v[r] = 0.0;
  if (STAGE >= 2) {
    if (inside) {
      const double2 q0 = *reinterpret_cast<const double2 *>(src);
      const double2 q1 = *reinterpret_cast<const double2 *>(src + 2);
      u0v[0] = q0.x; u0v[1] = q0.y; u0v[2] = q1.x; u0v[3] = q1.y;
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < left) u0v[r] = src[r];
    }
  }
}

// Ties an element offset to a value the compiler has only late in the kernel, so that the loads
// at that offset cannot be hoisted above that point (LATE variants: u0 is fetched after the
// reconstruction, which takes its 8 registers out of the most crowded part of the kernel).
// The offset, not the pointer, is laundered: the loads stay ld.global.
__device__ __forceinline__ int64_t after(int64_t off, double late) {
#ifndef PSK_HOST_EMU
  asm volatile("" : "+l"(off) : "d"(late));
#else
  (void)late;
#endif
  return off;
}

// ---------------------------------------------------------------------------
// stage_warp_fast_share_kernel (default layout of the specialised kernel; same arithmetic per cell
// as stage_warp_fast_kernel, hence bitwise the same results; psk_set_stage_variant(5000 + k) selects
// k = 0: stage_warp_fast_kernel, k = 2: this one):
// the 120-cell layout of stage_warp_fast_kernel (halo
// lanes 0 and 31), but every lane computes the first differences t and the second-difference
// squares pq of its OWN four cells only and gets the three t and two pq of its neighbours'
// cells by shuffle instead of recomputing them from shuffled cell values: 12 FP64-pipe
// instructions fewer per lane (237 -> 225) for 3 more double shuffles.
// LATE: u0 is fetched after the reconstruction instead of with the stage input -- one more exposed
// load latency (-6 % on the hot configuration, which therefore uses LATE = 0) but 8 registers fewer
// in the most crowded part of the kernel, which the other fluxes / equations need to stay free of
// spills at 64 registers.
template <int EQ, int FLUX, int STAGE, bool WITH_MAX, int LATE = (EQ == PSK_EQ_BURGERS && FLUX == PSK_FLUX_RUSANOV) ? 0 : 1>
__global__ void __launch_bounds__(256, PSK_FAST_MIN_BLOCKS)
stage_warp_fast_share_kernel(const FastParams p) {
  constexpr int R = 4;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chunk >= p.chunks_per_row) return;
  const int row = blockIdx.y + blockIdx.z * gridDim.y;
  const int c0 = chunk * 30 * R - R + R * lane;
  const bool inside = (c0 >= 0) && (c0 + R <= p.bc.n);
  const int g = p.bc.g, n = p.bc.n;
  const int64_t off = static_cast<int64_t>(row) * p.ld + g + c0;
  const bool emit = (lane >= 1) && (lane <= 30);
  FastIn in;
  fast_load<0>(p, row, c0, lane, inside, in);  // the stage input only
  double u0v[R] = {0.0, 0.0, 0.0, 0.0};
  if (LATE == 0 && emit) fast_load_u0<STAGE>(p.u0 + off, inside, n - c0, u0v);

  double w[R + 1];  // cells c0-1 .. c0+3
#pragma unroll
  for (int r = 0; r < R; ++r) w[1 + r] = in.v[r];
  w[0] = __shfl_up_sync(kFull, w[R], 1);
  // t[k]: sixth of the first difference over the interval (c0-2+k, c0-1+k), k = 0..6; own: k = 1..4
  double t[R + 3];
#pragma unroll
  for (int k = 1; k <= R; ++k) t[k] = __dmul_rn(1.0 / 6.0, w[k] - w[k - 1]);
  t[0] = __shfl_up_sync(kFull, t[R], 1);
  t[R + 1] = __shfl_down_sync(kFull, t[1], 1);
  t[R + 2] = __shfl_down_sync(kFull, t[2], 1);
  // pq[k]: centred at the cell c0-1+k, k = 0..5; own: k = 1..4
  double pq[R + 2];
#pragma unroll
  for (int k = 1; k <= R; ++k) {
    const double dd = t[k + 1] - t[k];
    pq[k] = fma((13.0 / 3.0) * dd, dd, p.eps9);
  }
  pq[0] = __shfl_up_sync(kFull, pq[R], 1);
  pq[R + 1] = __shfl_down_sync(kFull, pq[1], 1);

  double ul[R], ur[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const Weno5Pair o = weno53_pair_lean(w[1 + r], t[r], t[r + 1], t[r + 2], t[r + 3], pq[r], pq[r + 1], pq[r + 2]);
    ul[r] = o.ul;
    ur[r] = o.ur;
  }
  const double ur_left = __shfl_up_sync(kFull, ur[R - 1], 1);
  const double ul_right = __shfl_down_sync(kFull, ul[0], 1);
  if (LATE == 1 && emit) fast_load_u0<STAGE>(p.u0 + after(off, ur[R - 1]), inside, n - c0, u0v);

  // -2 s of the scaled Rusanov flux: -2 max(|a|, |b|) = the larger magnitude of -2|a|, -2|b|, one
  // DMUL per CELL (|.| is an operand modifier) instead of an FP64 abs per cell and a DMUL per face
  const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? -2.0 * p.lf_speed[row] : 0.0;
  double m2[R + 2];  // cells c0-1 .. c0+4, the last one from the lane to the right
  if (EQ == PSK_EQ_BURGERS && FLUX == PSK_FLUX_RUSANOV) {
#pragma unroll
    for (int j = 0; j <= R; ++j) m2[j] = -2.0 * fabs(w[j]);
    m2[R + 1] = __shfl_down_sync(kFull, m2[1], 1);
  }
  double F[R + 1];
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const double urj = (f == 0) ? ur_left : ur[f - 1];
    const double ulp = (f == R) ? ul_right : ul[f];
    if (EQ == PSK_EQ_BURGERS) {
      if (FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
        const double a2 = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? speed : umax_neg(m2[f], m2[f + 1]);
        F[f] = fma(a2, ulp - urj, fma(urj, urj, ulp * ulp));
      } else if (FLUX == PSK_FLUX_UPWIND) {
        const double x = (urj + ulp) > 0.0 ? urj : ulp;
        F[f] = __dmul_rn(x, x);  // never contracted with the flux difference (same bits in every kernel)
      } else {
        const double vp = fmax(urj, 0.0), vm = fmin(ulp, 0.0);
        F[f] = fma(vp, vp, vm * vm);
      }
    } else {
      const int j = g + c0 + f - 1;
      const bool ok = (j >= 0 && j < p.bc.nx - 1);
      const double arj = ok ? p.vel_r[j] : 0.0, alp = ok ? p.vel_l[j + 1] : 0.0;
      const bool pos = (arj + alp) > 0.0;
      F[f] = (EQ == PSK_EQ_ADVECTION) ? (pos ? urj : ulp) : (pos ? __dmul_rn(arj, urj) : __dmul_rn(alp, ulp));
    }
  }

  double coef = p.coef;
  if (STAGE != 0) coef *= p.dt[static_cast<int64_t>(row) * p.dt_stride];
  double out[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double dF = F[r] - F[r + 1];
    if (EQ == PSK_EQ_ADVECTION) dF *= (c0 + r >= 0 && c0 + r < n) ? p.vel[g + c0 + r] : 0.0;
    if (STAGE == 0) {
      out[r] = coef * dF;
    } else if (STAGE == 4) {
      out[r] = fma(p.cc * coef, dF, fma(p.ca, u0v[r], p.cb * w[1 + r]));
    } else {
      const double k = fma(coef, dF, w[1 + r]);
      out[r] = (STAGE == 1) ? k
                            : ((STAGE == 2) ? fma(0.25, k, 0.75 * u0v[r]) : fma(2.0 / 3.0, k, (1.0 / 3.0) * u0v[r]));
    }
  }
  if (emit) {
    if (inside) {
      *reinterpret_cast<double2 *>(p.uout + off) = make_double2(out[0], out[1]);
      *reinterpret_cast<double2 *>(p.uout + off + 2) = make_double2(out[2], out[3]);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (c0 + r >= 0 && c0 + r < n) p.uout[off + r] = out[r];
    }
  }
  if (WITH_MAX) {
    unsigned long long mx = 0ull;
    if (emit) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (c0 + r >= 0 && c0 + r < n) {
          const unsigned long long b = abs_bits(out[r]);
          mx = b > mx ? b : mx;
        }
    }
    mx = warp_max_bits(mx);
    if (lane == 0) atomicMax(p.maxabs + row, mx);
  }
}

// ---------------------------------------------------------------------------
// The whole SSPRK33 step of the hot configuration in ONE launch (temporal blocking over the three
// stages; timestepping.py:312-320 with the RHS of schemes.py:339-346): Burgers with the Rusanov,
// upwind ("Godunov") or Engquist-Osher flux, WENO-JS5 FAST, periodic rows.  u is read once and u' written once: 16 B of HBM traffic per
// cell-update instead of the 64 B of three stage launches, one load phase per step instead of
// three, and the stage values k1, k2 never leave the registers.
//
//   * a warp owns a window of 32 R cells of one row (lane l: the cells R l .. R l + R - 1, R even);
//     every stage is the shared-difference form of stage_warp_fast_share_kernel (own first
//     differences / second-difference squares / -2|w|, the neighbours' by shuffle), applied to the
//     stage values the lanes hold, so per cell the arithmetic is that of the stage kernels, bit
//     for bit;
//   * each stage invalidates 3 more cells at either end of the window (their stencils reach
//     outside it; lanes 0 and 31 shuffle with themselves there, the values are finite garbage that
//     never reaches a stored cell): after three stages the window cells [9, 32 R - 9) are valid,
//     of which the aligned range [10, 32 R - 10) is stored.  Consecutive windows of a row start
//     32 R - 20 cells apart;
//   * window cells beyond the row ends are the periodic images, or -- for a slab of a larger grid
//     (boundary kind NONE, g >= 9 ghost cells filled by the neighbours before the step) -- the row's
//     stored ghost cells (loaded through the slow path).
// FP64 work per emitted cell is that of the stage kernels (32 R / (32 R - 20) against 32 / 30).
struct StepParams {
  const double *u;
  double *uout;
  const double *dt;
  const uint8_t *active;  // rows with active[r] == 0 are copied through
  unsigned long long *maxabs;
  int64_t ld;
  double coef;  // 1 / (flux scale * dx): 4 for Rusanov, 2 for the upwind and Engquist-Osher fluxes
  double eps9;  // eps / 9
  int dt_stride;
  int chunks_per_row;
  int n, g;
  int bc_none;  // 1: window cells beyond the row ends are the row's stored ghost cells (slab of a larger grid, g >= 9)
  double *k1_out, *k2_out;  // STAGES kernels only: the stage values are stored too (uout may then be NULL)
  int shift;                // step_warp_fused_p2p_kernel only: the chunk grid starts this many cells left of the slab
  // DIRICHLET rows: boundary data at the three stage times t, t + dt, t + dt / 2 (timestepping.py:314-319) as three
  // consecutive blocks of ghost_block doubles; row r of a block at + r * ghost_ld (0: one set for all rows); 2 g
  // values per row, left ghost cells first
  const double *ghost3;
  int64_t ghost_ld, ghost_block;
  // advection / continuity: velocity and its reconstruction (time independent), nx entries each
  const double *vel, *vel_l, *vel_r;
  // Rusanov / Lax-Friedrichs with alpha != 1: the artificial viscosity nu = df ** (alpha - 1) of every face (j | j + 1)
  // of the ARRAY (scalar.py:231-234), nx - 1 entries; NU kernels only (rows with boundary data: on periodic rows the
  // face at the seam would need two values)
  const double *nu;
};

// what the upwind switch of the advection / continuity flux needs at the R + 1 faces of a lane's cells
// (advection/schemes.py:107-114, continuity/schemes.py:103-110): taken once per step, the velocity does not
// depend on time.  pos bit f: (ar + al) > 0 at the face between the cells own + f - 1 and own + f
template <int R>
struct StepVel {
  unsigned pos;
  double ar[R + 1], al[R + 1];  // continuity only
  double vc[R];                 // advection only: velocity of the own cells (0 outside the row)
};

template <int R, int EQ>
__device__ __forceinline__ void step_load_vel(const StepParams &p, int c0, bool periodic, StepVel<R> &v) {
  const int n = p.n, nx = p.n + 2 * p.g;
  // periodic rows: a window cell beyond the row ends is the image of an interior cell and is advanced with that
  // cell's velocity data (the caller guarantees that the velocity's reconstruction is periodic too, see
  // psk_ssprk33_step: then these are the very numbers the stage kernels read at the row ends)
  // Only the warps at a row end pay for the modulo: two copies of the loop behind one branch.
  auto load = [&](auto far) {
    auto wrap = [&](int c) {
      if (decltype(far)::value) {
        c %= n;
        if (c < 0) c += n;
      }
      return c;
    };
    v.pos = 0u;
#pragma unroll
    for (int f = 0; f <= R; ++f) {
      const int jl = p.g + wrap(c0 + f - 1), jr = p.g + wrap(c0 + f);  // array indices of the cells left / right of the face
      const bool ok = periodic || (jl >= 0 && jl < nx - 1);
      const double arj = ok ? p.vel_r[jl] : 0.0, alp = ok ? p.vel_l[jr] : 0.0;
      if ((arj + alp) > 0.0) v.pos |= 1u << f;
      v.ar[f] = arj;
      v.al[f] = alp;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = c0 + r;
      v.vc[r] = (EQ == PSK_EQ_ADVECTION && (periodic || (c >= 0 && c < n))) ? p.vel[p.g + wrap(c)] : 0.0;
    }
  };
  if (periodic && !(c0 - 1 >= 0 && c0 + R < n))
    load(std::true_type{});
  else
    load(std::false_type{});
}

// Dirichlet rows: the window cells that are ghost cells take the boundary data of `stage` (0, 1, 2); window
// cells beyond the ghost cells never reach a stored cell
template <int R>
__device__ __forceinline__ void step_fill_ghosts(const StepParams &p, int row, int stage, int c0, double (&a)[R]) {
  const double *gh = p.ghost3 + static_cast<int64_t>(stage) * p.ghost_block + static_cast<int64_t>(row) * p.ghost_ld;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = c0 + r;
    if (c < 0) a[r] = (c >= -p.g) ? gh[c + p.g] : 0.0;
    if (c >= p.n) a[r] = (c < p.n + p.g) ? gh[p.g + c - p.n] : 0.0;
  }
}

// Neumann rows (scalar.py:472-500): the ghost cell -1 - k takes the stage value of its mirror image k plus the
// boundary data of `stage`, n + k that of n - 1 - k (k = 0, 1, 2: a stage reaches no farther).  The image sits
// 2 k + 1 <= 5 < R places to the right (left end) / left (right end) in the window: in a register of the same lane or
// of its neighbour, both known at compile time -- five shuffles per side, then selects.  `left` / `right` (the window
// reaches beyond that end of the row) are warp-uniform.  Window cells beyond the three ghost cells never reach a
// stored cell.
template <int R>
__device__ __forceinline__ void step_fill_neumann(const StepParams &p, int row, int stage, int c0, bool left, bool right,
                                                  double (&a)[R]) {
  static_assert(R >= 6, "an image is at most five places away");
  constexpr unsigned kFull = 0xffffffffu;
  const double *gh = p.ghost3 + static_cast<int64_t>(stage) * p.ghost_block + static_cast<int64_t>(row) * p.ghost_ld;
  if (left) {
    double nb[5];  // the first five cells of the next lane
#pragma unroll
    for (int j = 0; j < 5; ++j) nb[j] = __shfl_down_sync(kFull, a[j], 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double src = (r + 2 * k + 1 < R) ? a[(r + 2 * k + 1) % R] : nb[(r + 2 * k + 1) % R % 5];
        if (c0 + r == -1 - k) a[r] = src + gh[p.g - 1 - k];
      }
    }
  }
  if (right) {
    double nb[5];  // the last five cells of the previous lane
#pragma unroll
    for (int j = 0; j < 5; ++j) nb[j] = __shfl_up_sync(kFull, a[R - 5 + j], 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double src = (r - 2 * k - 1 >= 0) ? a[(r - 2 * k - 1 + R) % R] : nb[(r - 2 * k - 1 + 5) % 5];
        if (c0 + r == p.n + k) a[r] = src + gh[p.g + k];
      }
    }
  }
}

// nu of the R + 1 faces of a lane's cells: face f lies between the cells c0 + f - 1 and c0 + f, i.e. it is face
// g + c0 + f - 1 of the array; 1 beyond the array (never reaches a stored cell)
template <int R>
__device__ __forceinline__ void step_load_nu(const StepParams &p, int c0, double (&nuf)[R + 1]) {
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const int j = p.g + c0 + f - 1;
    nuf[f] = (j >= 0 && j < p.n + 2 * p.g - 1) ? p.nu[j] : 1.0;
  }
}

template <int R>
struct StepGeometry {
  static constexpr int kWindow = 32 * R;
  static constexpr int kSkip = 10;                 // first stored window cell
  static constexpr int kEmit = kWindow - 2 * kSkip;  // stored cells per warp
};

// (no global wave speed: every flux but Lax-Friedrichs)
struct NoSpeed {
  __device__ __forceinline__ double operator()() const { return 0.0; }
};

// one stage on the lane's R cells: a (stage input, own cells) -> L = coef * dF per own cell.
// Lax-Friedrichs: `speed2()` returns -2 max |w| of the row (scalar.py:277) and is called AFTER the reconstruction,
// right before the fluxes -- the row-wide reduction it may have to wait for overlaps the bulk of the stage.
// NU: nuf[f] = nu of the face between the cells own + f - 1 and own + f multiplies the speed (scalar.py:231-249).
template <int R, int FLUX, int EQ = PSK_EQ_BURGERS, class SpeedFn = NoSpeed, bool NU = false>
__device__ __forceinline__ void step_stage_rhs(const double (&a)[R], double eps9, double (&dF)[R],
                                               const StepVel<R> *vel = nullptr, SpeedFn speed2 = SpeedFn(),
                                               const double *nuf = nullptr) {
  constexpr unsigned kFull = 0xffffffffu;
  double w0 = __shfl_up_sync(kFull, a[R - 1], 1);  // cell own - 1
  double t[R + 3];   // t[k]: interval (own - 2 + k, own - 1 + k); own: k = 1..R
  t[1] = __dmul_rn(1.0 / 6.0, a[0] - w0);
#pragma unroll
  for (int k = 2; k <= R; ++k) t[k] = __dmul_rn(1.0 / 6.0, a[k - 1] - a[k - 2]);
  t[0] = __shfl_up_sync(kFull, t[R], 1);
  t[R + 1] = __shfl_down_sync(kFull, t[1], 1);
  t[R + 2] = __shfl_down_sync(kFull, t[2], 1);
  double pq[R + 2];  // centred at the cell own - 1 + k; own: k = 1..R
#pragma unroll
  for (int k = 1; k <= R; ++k) {
    const double dd = t[k + 1] - t[k];
    pq[k] = fma((13.0 / 3.0) * dd, dd, eps9);
  }
  pq[0] = __shfl_up_sync(kFull, pq[R], 1);
  pq[R + 1] = __shfl_down_sync(kFull, pq[1], 1);
  double m2[R + 2];  // -2 |w| of the cells own - 1 .. own + R (Rusanov speed, as in the stage kernel)
  if (EQ == PSK_EQ_BURGERS && FLUX == PSK_FLUX_RUSANOV) {
    m2[0] = -2.0 * fabs(w0);
#pragma unroll
    for (int j = 1; j <= R; ++j) m2[j] = -2.0 * fabs(a[j - 1]);
    m2[R + 1] = __shfl_down_sync(kFull, m2[1], 1);
  }
  double ul[R], ur[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const Weno5Pair o = weno53_pair_lean(a[r], t[r], t[r + 1], t[r + 2], t[r + 3], pq[r], pq[r + 1], pq[r + 2]);
    ul[r] = o.ul;
    ur[r] = o.ur;
  }
  const double ur_left = __shfl_up_sync(kFull, ur[R - 1], 1);
  const double ul_right = __shfl_down_sync(kFull, ul[0], 1);
  double lf2 = 0.0;
  if (EQ == PSK_EQ_BURGERS && FLUX == PSK_FLUX_LAX_FRIEDRICHS) lf2 = speed2();
  double F[R + 1];
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const double urj = (f == 0) ? ur_left : ur[f - 1];
    const double ulp = (f == R) ? ul_right : ul[f];
    if (EQ != PSK_EQ_BURGERS) {  // upwind by the sign of the averaged velocity, as in the stage kernels
      const bool pos = (vel->pos >> f) & 1u;
      F[f] = (EQ == PSK_EQ_ADVECTION) ? (pos ? urj : ulp) : (pos ? __dmul_rn(vel->ar[f], urj) : __dmul_rn(vel->al[f], ulp));
    } else if (FLUX == PSK_FLUX_RUSANOV) {  // 4 F (scalar.py:231-249)
      double a2 = umax_neg(m2[f], m2[f + 1]);
      if (NU) a2 = __dmul_rn(nuf[f], a2);  // -2 (nu a): the bits of `a *= nu` in the stage kernels
      F[f] = fma(a2, ulp - urj, fma(urj, urj, ulp * ulp));
    } else if (FLUX == PSK_FLUX_LAX_FRIEDRICHS) {  // 4 F (scalar.py:258-278), as in the stage kernels
      const double a2 = NU ? __dmul_rn(nuf[f], lf2) : lf2;
      F[f] = fma(a2, ulp - urj, fma(urj, urj, ulp * ulp));
    } else if (FLUX == PSK_FLUX_UPWIND) {  // 2 F (scalar.py:123-132)
      const double x = (urj + ulp) > 0.0 ? urj : ulp;
      F[f] = __dmul_rn(x, x);  // never contracted with the flux difference (same bits in every kernel)
    } else {  // Engquist-Osher, omega = 0: 2 F (scalar.py:311-322)
      const double vp = fmax(urj, 0.0), vm = fmin(ulp, 0.0);
      F[f] = fma(vp, vp, vm * vm);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    dF[r] = F[r] - F[r + 1];
    if (EQ == PSK_EQ_ADVECTION) dF[r] *= vel->vc[r];  // non-conservative form (advection/schemes.py:73)
  }
}

// a[0..R) -> dst on the stored cells of the lane (aligned pairs where the whole quad lies in the row)
template <int R>
__device__ __forceinline__ void step_store(double *dst, bool inside, const bool (&st)[R], const double (&a)[R]) {
  if (inside) {
#pragma unroll
    for (int r = 0; r < R; r += 2)
      if (st[r]) *reinterpret_cast<double2 *>(dst + r) = make_double2(a[r], a[r + 1]);
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st[r]) dst[r] = a[r];
  }
}

// STAGES: the stage values k1, k2 are stored as well (they are valid on more than the stored range), and
// the third stage is skipped when no uout is given: the recomputation of the reverse sweep, which needs
// k1 and k2 of a checkpointed state (and the next state, inside a tape segment), in one launch
// instead of two or three.
// EQ / BCK: the same kernel for the advection and continuity equations (upwind flux with the velocity's
// reconstruction) and for rows with boundary data, BCK = 1 Dirichlet, 2 Neumann -- the window cells that are ghost
// cells take the data of each stage time (Neumann: plus the stage value of their mirror image) before that stage,
// exactly where apply_boundary writes them (scalar.py:418-427, 472-500).
template <int R, int FLUX, bool WITH_MAX, int THREADS, int MINB, bool STAGES = false, int EQ = PSK_EQ_BURGERS,
          int BCK = 0, bool NU = false>
__global__ void __launch_bounds__(THREADS, MINB)
step_warp_fused_kernel(const StepParams p) {
  using Geo = StepGeometry<R>;
  static_assert(R % 2 == 0, "aligned pairs");
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chunk >= p.chunks_per_row) return;
  const int row = blockIdx.y + blockIdx.z * gridDim.y;
  const int n = p.n;
  constexpr bool DIRICHLET = BCK == 1, NEUMANN = BCK == 2;
  const int wstart = chunk * Geo::kEmit - Geo::kSkip;  // first window cell (interior coordinates)
  const int c0 = wstart + R * lane;                    // first own cell
  const bool inside = (c0 >= 0) && (c0 + R <= n);
  const bool edge = (wstart < 0) || (wstart + Geo::kWindow > n);  // the window reaches beyond the row (warp-uniform)
  const int64_t base = static_cast<int64_t>(row) * p.ld + p.g;
  // which own cells are stored: window cells [kSkip, kWindow - kSkip) that exist in the row
  const int wc0 = R * lane;
  bool st[R];
#pragma unroll
  for (int r = 0; r < R; ++r)
    st[r] = (wc0 + r >= Geo::kSkip) && (wc0 + r < Geo::kWindow - Geo::kSkip) && (c0 + r >= 0) && (c0 + r < n);

  double u0[R];
  if (inside) {
#pragma unroll
    for (int r = 0; r < R; r += 2) {
      const double2 q = *reinterpret_cast<const double2 *>(p.u + base + c0 + r);
      u0[r] = q.x;
      u0[r + 1] = q.y;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int c = c0 + r;
      if (NEUMANN) {  // apply_boundary straight from the stored state (scalar.py:472-500)
        const double *gh = p.ghost3 + static_cast<int64_t>(row) * p.ghost_ld;
        u0[r] = (c >= 0 && c < n)        ? p.u[base + c]
                : (c >= -3 && c < 0)     ? p.u[base - 1 - c] + gh[p.g + c]
                : (c >= n && c < n + 3)  ? p.u[base + 2 * n - 1 - c] + gh[p.g + c - n]
                                         : 0.0;
      } else if (DIRICHLET) {
        u0[r] = (c >= 0 && c < n) ? p.u[base + c] : 0.0;  // ghost cells: step_fill_ghosts below
      } else if (p.bc_none) {  // ghost cells filled by the neighbouring slabs; further out: never reaches a stored cell
        u0[r] = (c >= -p.g && c < n + p.g) ? p.u[base + c] : 0.0;
      } else {
        c %= n;  // periodic image (n >= 1)
        if (c < 0) c += n;
        u0[r] = p.u[base + c];
      }
    }
  }
  if (p.active != nullptr && p.active[row] == 0) {  // finished row: the state is carried over
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st[r]) p.uout[base + c0 + r] = u0[r];
    return;
  }
  const double cdt = p.coef * p.dt[static_cast<int64_t>(row) * p.dt_stride];
  StepVel<R> vel;
  if (EQ != PSK_EQ_BURGERS) step_load_vel<R, EQ>(p, c0, BCK == 0 && !p.bc_none, vel);
  double nuf[R + 1];
  if constexpr (NU) step_load_nu<R>(p, c0, nuf);
  if (DIRICHLET && !inside) step_fill_ghosts<R>(p, row, 0, c0, u0);

  double a[R], dF[R];
  step_stage_rhs<R, FLUX, EQ, NoSpeed, NU>(u0, p.eps9, dF, &vel, NoSpeed(), nuf);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(cdt, dF[r], u0[r]);  // k1
  if (STAGES) step_store<R>(p.k1_out + base + c0, inside, st, a);
  if (DIRICHLET && !inside) step_fill_ghosts<R>(p, row, 1, c0, a);
  if constexpr (NEUMANN) {
    if (edge) step_fill_neumann<R>(p, row, 1, c0, wstart < 0, wstart + Geo::kWindow > n, a);
  }
  step_stage_rhs<R, FLUX, EQ, NoSpeed, NU>(a, p.eps9, dF, &vel, NoSpeed(), nuf);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(0.25, fma(cdt, dF[r], a[r]), 0.75 * u0[r]);  // k2
  if (STAGES) {
    step_store<R>(p.k2_out + base + c0, inside, st, a);
    if (p.uout == nullptr) return;
  }
  if (DIRICHLET && !inside) step_fill_ghosts<R>(p, row, 2, c0, a);
  if constexpr (NEUMANN) {
    if (edge) step_fill_neumann<R>(p, row, 2, c0, wstart < 0, wstart + Geo::kWindow > n, a);
  }
  step_stage_rhs<R, FLUX, EQ, NoSpeed, NU>(a, p.eps9, dF, &vel, NoSpeed(), nuf);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(2.0 / 3.0, fma(cdt, dF[r], a[r]), (1.0 / 3.0) * u0[r]);  // u'

  if (inside) {
#pragma unroll
    for (int r = 0; r < R; r += 2)
      if (st[r]) *reinterpret_cast<double2 *>(p.uout + base + c0 + r) = make_double2(a[r], a[r + 1]);
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st[r]) p.uout[base + c0 + r] = a[r];
  }
  if (WITH_MAX) {
    unsigned long long mx = 0ull;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st[r]) {
        const unsigned long long b = abs_bits(a[r]);
        mx = b > mx ? b : mx;
      }
    mx = warp_max_bits(mx);
    if (lane == 0) atomicMax(p.maxabs + row, mx);
  }
}

}  // namespace psk
