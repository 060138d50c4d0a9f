// psk_forward.cu -- forward hot path of libpsk: apply_boundary + WENO-JS reconstruction +
// numerical flux + flux-difference RHS fused with the SSPRK33 stage combine, the CFL
// max-wavespeed reduction, and the small parity entry points (reconstruct, numerical_flux,
// apply_boundary).  Hand-written fp64 CUDA for sm_100a; no tensor cores (nothing here is a
// contraction), no library calls.
//
// Tile kernel (stage_tile_kernel)
//   * one CTA owns TILE = R * blockDim.x consecutive INTERIOR cells of one row;
//   * the tile plus its 3-cell halo is staged once in shared memory with coalesced loads,
//     the boundary condition being applied on the fly (ghost cells are never read from
//     the state array for periodic / Dirichlet / Neumann rows);
//   * every thread then owns R consecutive cells: first/second differences, smoothness
//     indicators, nonlinear weights, fluxes and the stage update live in registers, so a
//     stage reads u (and u0) once and writes once;
//   * neighbouring threads exchange one left and one right face value through shared
//     memory (a single __syncthreads) instead of recomputing them;
//   * the shared tile is padded by one double every R entries so that the stride-R
//     per-thread window reads are bank-conflict free.
// Ghost ROWS of the output (which the reference also produces, from zero-padded stencils)
// are written by a separate tiny kernel only on request.
#include <cuda_runtime.h>

#include <cstdint>

#include "psk_common.cuh"
#include "psk_math.cuh"
#include "psk_fast_kernels.cuh"

namespace psk {

thread_local int g_last_cuda_error = 0;

struct StageParams {
  const double *uin;   // state the RHS is evaluated on (k_{s-1})
  const double *u0;    // state at the start of the step
  double *uout;
  const double *dt;
  int64_t dt_stride;
  const uint8_t *active;
  const double *lf_speed;
  unsigned long long *maxabs;
  const double *nu;     // per-face dissipation scale or nullptr
  const double *vel;    // advection / continuity
  const double *vel_l;
  const double *vel_r;
  BcView bc;
  int64_t ld;
  double dx, invdx, eps;
  double delta;  // ESWENO32
  int stage;
  int tiles_per_row;
  int vec_ok;  // rows are 16-byte aligned and every tile starts on an even cell
  double ca, cb, cc;  // stage 4 (psk_rhs_axpby): uout = ca u0 + cb uin + cc dt L(uin)
};

// stage_combine (psk_math.cuh) plus stage 4, the general combine the other Runge-Kutta methods are built from
// (timestepping.py:289-405: ForwardEuler, RK44, CKRK45)
template <bool STRICT>
__device__ __forceinline__ double stage_combine_p(const StageParams &p, double u0, double w, double dt, double L) {
  if (p.stage == 4) {
    if (STRICT) return sadd(sadd(smul(p.ca, u0), smul(p.cb, w)), smul(p.cc, smul(dt, L)));
    return fma(p.cc * dt, L, fma(p.ca, u0, p.cb * w));
  }
  return stage_combine<STRICT>(p.stage, u0, w, dt, L);
}

template <int R>
__host__ __device__ __forceinline__ int pad_index(int e) {
  return e + e / R;
}

template <int EQ, int FLUX, int REC, bool STRICT, int R>
__global__ void __launch_bounds__(256)
stage_tile_kernel(const StageParams p) {
  extern __shared__ double smem[];
  const int nthreads = blockDim.x;
  const int tile_cells = R * nthreads;
  const int row = blockIdx.x / p.tiles_per_row;
  const int tile = blockIdx.x - row * p.tiles_per_row;
  if (p.active != nullptr && p.active[row] == 0) return;

  const int g = p.bc.g;
  const int n = p.bc.n;
  const int s = tile * tile_cells;  // first interior cell of the tile (interior coordinates)
  const double *__restrict__ urow = p.uin + static_cast<int64_t>(row) * p.ld;

  // shared layout: padded tile | XL[nthreads + 1] | XR[nthreads + 1] | warp maxima
  double *tile_w = smem;
  const int tile_elems = tile_cells + 2 * kHalo;
  double *xl = tile_w + pad_index<R>(tile_elems) + 1;
  double *xr = xl + nthreads + 1;

  // ---- stage the tile: element e <-> array index g + s - kHalo + e
  for (int e = threadIdx.x; e < tile_elems; e += nthreads)
    tile_w[pad_index<R>(e)] = load_w(p.bc, urow, row, g + s - kHalo + e);
  __syncthreads();

  // ---- per-thread window: cells c0 - 3 .. c0 + R + 2 (interior coordinates)
  const int t = threadIdx.x;
  const int c0 = s + R * t;
  double v[R + 2 * kHalo];
  {
    const double *base = tile_w + (R + 1) * t;
#pragma unroll
    for (int k = 0; k < R + 2 * kHalo; ++k) v[k] = base[k + k / R];
  }

  // ---- face values of the R owned cells
  double ul[R], ur[R];
  if (REC == PSK_REC_WENOJS53 && !STRICT) {
    // half first differences hd[k] = 0.5 (v[k+1] - v[k]), and 13/12 (second difference)^2
    double hd[R + 5], pq[R + 4];
#pragma unroll
    for (int k = 0; k < R + 5; ++k) hd[k] = 0.5 * (v[k + 1] - v[k]);
#pragma unroll
    for (int k = 0; k < R + 4; ++k) {
      double tt = hd[k + 1] - hd[k];  // centred at v[k + 1]
      pq[k] = (13.0 / 3.0) * tt * tt;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int m = r + kHalo;  // window index of the cell
      Weno5Pair o = weno53_pair_fast(v[m], hd[m - 2], hd[m - 1], hd[m], hd[m + 1], pq[m - 2],
                                     pq[m - 1], pq[m], p.eps);
      ul[r] = o.ul;
      ur[r] = o.ur;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int m = r + kHalo;
      Weno5Pair o =
          reconstruct_cell<REC, STRICT>(v[m - 2], v[m - 1], v[m], v[m + 1], v[m + 2], p.eps);
      ul[r] = o.ul;
      ur[r] = o.ur;
    }
  }

  // ---- exchange: ur of the cell left of c0 and ul of the cell right of c0 + R - 1
  xl[t] = ul[0];
  xr[t + 1] = ur[R - 1];
  if (t == 0) {
    Weno5Pair o = reconstruct_cell<REC, STRICT>(v[0], v[1], v[2], v[3], v[4], p.eps);
    xr[0] = o.ur;  // cell c0 - 1
  }
  if (t == nthreads - 1) {
    Weno5Pair o = reconstruct_cell<REC, STRICT>(v[R + 1], v[R + 2], v[R + 3], v[R + 4], v[R + 5],
                                                 p.eps);
    xl[nthreads] = o.ul;  // cell c0 + R
  }
  __syncthreads();
  const double ur_left = xr[t];
  const double ul_right = xl[t + 1];

  // ---- fluxes at the R + 1 faces of the owned cells; face f sits between cells c0+f-1, c0+f
  const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? p.lf_speed[row] : 0.0;
  double F[R + 1];
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const double urj = (f == 0) ? ur_left : ur[f - 1];
    const double ulp = (f == R) ? ul_right : ul[f];
    const int j = g + c0 + f - 1;  // array index of the cell left of the face
    double nu = 1.0, arj = 0.0, alp = 0.0;
    if ((FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) && p.nu != nullptr)
      nu = (j >= 0 && j < p.bc.nx - 1) ? p.nu[j] : 1.0;
    if (EQ != PSK_EQ_BURGERS) {
      const bool ok = (j >= 0 && j < p.bc.nx - 1);
      arj = ok ? p.vel_r[j] : 0.0;
      alp = ok ? p.vel_l[j + 1] : 0.0;
    }
    F[f] = face_flux<EQ, FLUX, STRICT>(urj, ulp, v[f + kHalo - 1], v[f + kHalo], speed, nu, arj,
                                       alp);
  }

  // ---- RHS, stage combine, store
  const double dt = (p.stage != 0) ? p.dt[static_cast<int64_t>(row) * p.dt_stride] : 0.0;
  const int64_t off = static_cast<int64_t>(row) * p.ld + g + c0;
  double out[R];
  unsigned long long mx = 0ull;
  const bool full = (c0 + R <= n);
  double u0v[R];
  if (p.stage >= 2) {
    if (full && p.vec_ok && (R % 2 == 0)) {
#pragma unroll
      for (int r = 0; r < R; r += 2) {
        double2 q = *reinterpret_cast<const double2 *>(p.u0 + off + r);
        u0v[r] = q.x;
        u0v[r + 1] = q.y;
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) u0v[r] = (c0 + r < n) ? p.u0[off + r] : 0.0;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) u0v[r] = 0.0;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double vel = 0.0;
    if (EQ == PSK_EQ_ADVECTION) vel = (c0 + r < n) ? p.vel[g + c0 + r] : 0.0;
    const double L = rhs_from_faces<EQ, STRICT>(F[r], F[r + 1], vel, p.dx, p.invdx);
    out[r] = stage_combine_p<STRICT>(p, u0v[r], v[r + kHalo], dt, L);
    if (c0 + r < n) {
      const unsigned long long b = abs_bits(out[r]);
      mx = b > mx ? b : mx;
    }
  }
  if (full && p.vec_ok && (R % 2 == 0)) {
#pragma unroll
    for (int r = 0; r < R; r += 2)
      *reinterpret_cast<double2 *>(p.uout + off + r) = make_double2(out[r], out[r + 1]);
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (c0 + r < n) p.uout[off + r] = out[r];
  }

  // ---- fused CFL reduction: max |uout| over the interior of the row
  if (p.maxabs != nullptr) {
    mx = warp_max_bits(mx);
    unsigned long long *wm = reinterpret_cast<unsigned long long *>(xr + nthreads + 1);
    const int lane = t & 31, wid = t >> 5;
    if (lane == 0) wm[wid] = mx;
    __syncthreads();
    if (wid == 0) {
      const int nw = (nthreads + 31) >> 5;
      unsigned long long m2 = lane < nw ? wm[lane] : 0ull;
      m2 = warp_max_bits(m2);
      if (lane == 0) atomicMax(p.maxabs + row, m2);
    }
  }
}

// ---------------------------------------------------------------------------
// Warp kernel (stage_warp_kernel): the same stage without shared memory or block barriers.
//   * one WARP covers 32 * R consecutive cells of one row; every lane loads its R cells
//     with 128-bit loads straight into registers (one fully coalesced kilobyte per warp)
//     and takes its 3-cell halo from the neighbouring lanes with shuffles;
//   * lanes 0 and 31 are HALO lanes: they load and reconstruct like every other lane
//     (uniform code, no divergent extra work) but store nothing, so a warp emits
//     30 * R = 120 cells; the left / right face values of neighbouring cells also travel
//     by shuffle;
//   * no __syncthreads anywhere, so warps drift apart and hide each other's load latency;
//   * FAST math works in sixths of the first differences and with fluxes scaled by a
//     compile-time constant (folded into dt / dx), see psk_math.cuh.

template <int EQ, int FLUX>
__device__ __forceinline__ double face_flux_scaled(double urj, double ulp, double wj, double wp,
                                                   double speed, double nu, bool has_nu, double arj,
                                                   double alp) {
  if (EQ == PSK_EQ_BURGERS) {
    if (FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
      // 4 F = ul^2 + ur^2 - 2 a nu (ul - ur)      (scalar.py:231-249)
      double a = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? speed : fmax(fabs(wp), fabs(wj));
      if (has_nu) a *= nu;
      return fma(-2.0 * a, ulp - urj, fma(urj, urj, ulp * ulp));
    }
    if (FLUX == PSK_FLUX_UPWIND || FLUX == PSK_FLUX_ESWENO) {
      const double v = (urj + ulp) > 0.0 ? urj : ulp;  // scalar.py:129-130
      return __dmul_rn(v, v);  // never contracted with the flux difference (same bits in every kernel)
    }
    const double vp = fmax(urj, 0.0), vm = fmin(ulp, 0.0);  // scalar.py:312-321, omega = 0
    return fma(vp, vp, vm * vm);
  }
  const bool pos = (arj + alp) > 0.0;
  if (EQ == PSK_EQ_ADVECTION) return pos ? urj : ulp;
  return pos ? __dmul_rn(arj, urj) : __dmul_rn(alp, ulp);
}

// where a lane's R cells live and how they may be accessed
struct ChunkRef {
  int row, c0;
  int64_t off;
  bool vec, emit, live;
};

template <int R>
__device__ __forceinline__ ChunkRef chunk_ref(const StageParams &p, long long warp_id, int lane) {
  ChunkRef c;
  c.row = static_cast<int>(warp_id / p.tiles_per_row);
  const int chunk = static_cast<int>(warp_id - static_cast<long long>(c.row) * p.tiles_per_row);
  c.c0 = chunk * (30 * R) - R + R * lane;  // first owned cell (interior coordinates)
  c.off = static_cast<int64_t>(c.row) * p.ld + p.bc.g + c.c0;
  c.vec = (c.c0 >= 0) && (c.c0 + R <= p.bc.n) && p.vec_ok;
  c.emit = (lane >= 1) && (lane <= 30);
  c.live = (p.active == nullptr) || (p.active[c.row] != 0);
  return c;
}

// issue the global loads of one chunk: the lane's R cells of uin (boundary condition applied
// where the lane touches ghost cells) and, for stages 2 and 3, of u0
template <int R>
__device__ __forceinline__ void chunk_load(const StageParams &p, const ChunkRef &c, double (&own)[R],
                                           double (&u0v)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) u0v[r] = 0.0;
  if (!c.live) {
#pragma unroll
    for (int r = 0; r < R; ++r) own[r] = 0.0;
    return;
  }
  if (c.vec) {
    const double2 q0 = *reinterpret_cast<const double2 *>(p.uin + c.off);
    const double2 q1 = *reinterpret_cast<const double2 *>(p.uin + c.off + 2);
    own[0] = q0.x; own[1] = q0.y; own[2] = q1.x; own[3] = q1.y;
  } else {
    const double *__restrict__ urow = p.uin + static_cast<int64_t>(c.row) * p.ld;
#pragma unroll
    for (int r = 0; r < R; ++r) own[r] = load_w(p.bc, urow, c.row, p.bc.g + c.c0 + r);
  }
  if (p.stage >= 2 && c.emit) {
    if (c.vec) {
      const double2 q0 = *reinterpret_cast<const double2 *>(p.u0 + c.off);
      const double2 q1 = *reinterpret_cast<const double2 *>(p.u0 + c.off + 2);
      u0v[0] = q0.x; u0v[1] = q0.y; u0v[2] = q1.x; u0v[3] = q1.y;
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (c.c0 + r >= 0 && c.c0 + r < p.bc.n) u0v[r] = p.u0[c.off + r];
    }
  }
}

template <int EQ, int FLUX, int REC, bool STRICT, int R>
__device__ __forceinline__ void chunk_compute(const StageParams &p, const ChunkRef &c, int lane,
                                              const double (&own)[R], const double (&u0v)[R]) {
  static_assert(R == 4, "the shuffle pattern below is written for R = 4");
  constexpr unsigned kFull = 0xffffffffu;
  const int g = p.bc.g, n = p.bc.n;
  const int row = c.row, c0 = c.c0;
  const int64_t off = c.off;
  const bool vec = c.vec, emit = c.emit;
  double v[R + 2 * kHalo];
#pragma unroll
  for (int r = 0; r < R; ++r) v[kHalo + r] = own[r];
  v[0] = __shfl_up_sync(kFull, v[4], 1);
  v[1] = __shfl_up_sync(kFull, v[5], 1);
  v[2] = __shfl_up_sync(kFull, v[6], 1);
  v[7] = __shfl_down_sync(kFull, v[3], 1);
  v[8] = __shfl_down_sync(kFull, v[4], 1);
  v[9] = __shfl_down_sync(kFull, v[5], 1);
  // (lane 0 keeps its own values in v[0..2], lane 31 in v[7..9]: finite, and never used
  //  for anything that is stored)

  // ---- face values of the owned cells
  double ul[R], ur[R];
  if (REC == PSK_REC_WENOJS53 && !STRICT) {
    double t[R + 5], pq[R + 4];
    const double eps9 = p.eps * (1.0 / 9.0);
#pragma unroll
    for (int k = 0; k < R + 5; ++k) t[k] = __dmul_rn(1.0 / 6.0, v[k + 1] - v[k]);
#pragma unroll
    for (int k = 0; k < R + 4; ++k) {
      const double dd = t[k + 1] - t[k];  // centred at v[k + 1]
      pq[k] = fma((13.0 / 3.0) * dd, dd, eps9);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int m = r + kHalo;
      // the same function as the specialised kernels: the two paths give the same bits
      const Weno5Pair o = weno53_pair_lean(v[m], t[m - 2], t[m - 1], t[m], t[m + 1], pq[m - 2], pq[m - 1], pq[m]);
      ul[r] = o.ul;
      ur[r] = o.ur;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int m = r + kHalo;
      const Weno5Pair o =
          reconstruct_cell<REC, STRICT>(v[m - 2], v[m - 1], v[m], v[m + 1], v[m + 2], p.eps);
      ul[r] = o.ul;
      ur[r] = o.ur;
    }
  }
  const double ur_left = __shfl_up_sync(kFull, ur[R - 1], 1);
  const double ul_right = __shfl_down_sync(kFull, ul[0], 1);
  // ESWENO32 scheme: omega_0 of cells c0 - 1 .. c0 + R (burgers/schemes.py:237-240)
  double om[R + 2];
  if (FLUX == PSK_FLUX_ESWENO) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int m = r + kHalo;
      om[r + 1] = esweno32_cell<STRICT>(v[m - 1], v[m], v[m + 1], p.eps, false).om0;
    }
    om[0] = __shfl_up_sync(kFull, om[R], 1);
    om[R + 1] = __shfl_down_sync(kFull, om[1], 1);
  }

  // ---- fluxes at the R + 1 faces; face f sits between cells c0 + f - 1 and c0 + f
  const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? p.lf_speed[row] : 0.0;
  const bool has_nu = (FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) && (p.nu != nullptr);
  double F[R + 1];
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const double urj = (f == 0) ? ur_left : ur[f - 1];
    const double ulp = (f == R) ? ul_right : ul[f];
    double nu = 1.0, arj = 0.0, alp = 0.0;
    if (has_nu || EQ != PSK_EQ_BURGERS) {
      const int j = g + c0 + f - 1;  // array index of the cell left of the face
      const bool ok = (j >= 0 && j < p.bc.nx - 1);
      if (has_nu) nu = ok ? p.nu[j] : 1.0;
      if (EQ != PSK_EQ_BURGERS) {
        arj = ok ? p.vel_r[j] : 0.0;
        alp = ok ? p.vel_l[j + 1] : 0.0;
      }
    }
    if (STRICT)
      F[f] = face_flux<EQ, FLUX, true>(urj, ulp, v[f + kHalo - 1], v[f + kHalo], speed, nu, arj, alp);
    else
      F[f] = face_flux_scaled<EQ, FLUX>(urj, ulp, v[f + kHalo - 1], v[f + kHalo], speed, nu, has_nu,
                                        arj, alp);
    if (FLUX == PSK_FLUX_ESWENO) {
      const double gn = esweno_gnum<STRICT>(om[f], om[f + 1], v[f + kHalo - 1], v[f + kHalo], p.delta);
      F[f] = STRICT ? sadd(F[f], gn) : F[f] + gn;
    }
  }

  // ---- RHS, stage combine, store
  const double dt = (p.stage != 0) ? p.dt[static_cast<int64_t>(row) * p.dt_stride] : 0.0;
  double out[R];
  if (STRICT) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double vel = 0.0;
      if (EQ == PSK_EQ_ADVECTION) vel = (c0 + r >= 0 && c0 + r < n) ? p.vel[g + c0 + r] : 0.0;
      const double L = rhs_from_faces<EQ, true>(F[r], F[r + 1], vel, p.dx, p.invdx);
      out[r] = stage_combine_p<true>(p, u0v[r], v[r + kHalo], dt, L);
    }
  } else {
    const double cL = p.invdx * (1.0 / FluxScale<EQ, FLUX>::value);
    const double coef = (p.stage == 0) ? cL : dt * cL;  // dt / (scale dx)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double dF = F[r] - F[r + 1];
      if (EQ == PSK_EQ_ADVECTION) dF *= (c0 + r >= 0 && c0 + r < n) ? p.vel[g + c0 + r] : 0.0;
      if (p.stage == 0) {
        out[r] = coef * dF;
      } else if (p.stage == 4) {
        out[r] = fma(p.cc * coef, dF, fma(p.ca, u0v[r], p.cb * v[r + kHalo]));
      } else {
        const double k = fma(coef, dF, v[r + kHalo]);
        out[r] = (p.stage == 1) ? k
                                : ((p.stage == 2) ? fma(0.25, k, 0.75 * u0v[r])
                                                  : fma(2.0 / 3.0, k, (1.0 / 3.0) * u0v[r]));
      }
    }
  }
  if (emit) {
    if (vec) {
      *reinterpret_cast<double2 *>(p.uout + off) = make_double2(out[0], out[1]);
      *reinterpret_cast<double2 *>(p.uout + off + 2) = make_double2(out[2], out[3]);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (c0 + r >= 0 && c0 + r < n) p.uout[off + r] = out[r];
    }
  }

  // ---- fused CFL reduction (only when asked for): max |uout| over the interior of the row
  if (p.maxabs != nullptr) {
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

static int g_fast_wpc_max = 4;      // warps per CTA of the specialised stage kernel: 4 beats 7-8 by 3.5 % (a CTA waits for its slowest warp)
static int g_fast_layout = 2;   // 0: neighbours' differences recomputed from shuffled cells, 2: shared by shuffle (default)

// ---------------------------------------------------------------------------
// The specialised stage kernel FUSED with the ghost-cell exchange of a slab-decomposed grid
// (one row per GPU, boundary kind NONE, Burgers + Rusanov): one launch per stage and no other
// kernel, stream or event on the exchange path.
//   * the warps whose windows reach into the ghost cells (chunk 0; the chunks covering cells
//     >= n - 4) spin on the LOCAL epoch flags their neighbours raise -- every other warp of the
//     grid starts at once, so the NVLink latency hides behind the interior of the same launch;
//   * ghost cells are read with ld.volatile (they were written by another GPU while this
//     kernel may already have been running);
//   * the lane that stores cells 0..2 also stores them into the LEFT neighbour's right ghost
//     slots, the lane that stores cells n-3..n-1 into the RIGHT neighbour's left ghost slots
//     (peer pointers over NVLink), each followed by __threadfence_system() and a release store
//     of epoch + 1 to that neighbour's flag.
// A neighbour's ghost slots of `uout` are free to be overwritten: it last read them two
// stages ago, and this warp has just seen its push of the previous stage.
struct HaloLink {
  const long long *wait_lo, *wait_hi;
  long long wait_epoch;
  double *peer_lo, *peer_hi;
  long long *flag_lo, *flag_hi;
  unsigned long long timeout_ns;
  int *timed_out;
  const long long *epoch_in;  // optional device-side epoch (CUDA-graph replay)
  long long *epoch_out;
};

// (not inlined: the spin loop, its timer and its sleep stay out of the register allocation of the kernels)
__device__ __noinline__ void halo_spin(const long long *flag, long long epoch, unsigned long long timeout_ns,
                                       int *timed_out) {
  unsigned long long t0 = 0ull;
  bool timing = false;
  while (true) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= epoch) return;
    // sticky: once a wait has given up, no later wait spends its timeout again
    if (timed_out != nullptr && *reinterpret_cast<volatile int *>(timed_out) != 0) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (!timing) {
      t0 = now;
      timing = true;
    } else if (now - t0 > timeout_ns) {
      if (timed_out != nullptr) atomicExch(timed_out, 1);
      return;
    }
    __nanosleep(32);
  }
}

template <int STAGE, bool WITH_MAX>
__global__ void __launch_bounds__(256, PSK_FAST_MIN_BLOCKS)
stage_warp_fast_p2p_kernel(const FastParams p, const HaloLink h) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chunk >= p.chunks_per_row) return;
  const int g = p.bc.g, n = p.bc.n, nx = p.bc.nx;
  const int c0 = chunk * 30 * R - R + R * lane;
  const bool inside = (c0 >= 0) && (c0 + R <= n);
  const bool emit = (lane >= 1) && (lane <= 30);
  const bool edge_lo = (chunk == 0), edge_hi = (chunk * 30 * R + 31 * R > n);
  long long epoch = h.wait_epoch;
  // (the two pushing lanes live in chunk 0 and in the chunk holding cell n - 4: both edge chunks)
  if (h.epoch_in != nullptr && (edge_lo || edge_hi))
    epoch += *reinterpret_cast<const volatile long long *>(h.epoch_in);
  if (edge_lo && h.wait_lo != nullptr) halo_spin(h.wait_lo, epoch, h.timeout_ns, h.timed_out);
  if (edge_hi && h.wait_hi != nullptr) halo_spin(h.wait_hi, epoch, h.timeout_ns, h.timed_out);
  FastIn in;
  if (inside) {
    fast_load<STAGE>(p, 0, c0, lane, true, in);
  } else {
    const volatile double *urow = p.uin;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = g + c0 + r;
      in.v[r] = (i >= 0 && i < nx) ? urow[i] : 0.0;
      in.u0[r] = 0.0;
      if (STAGE >= 2 && emit && c0 + r >= 0 && c0 + r < n) in.u0[r] = p.u0[i];
    }
  }
  double out[4];
  fast_compute_store<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV, STAGE, WITH_MAX>(p, 0, c0, lane, inside, in, out);
  if (emit && inside) {
    if (c0 == 0 && h.peer_lo != nullptr) {
      h.peer_lo[0] = out[0];
      h.peer_lo[1] = out[1];
      h.peer_lo[2] = out[2];
      __threadfence_system();
      asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(h.flag_lo), "l"(epoch + 1) : "memory");
    }
    if (c0 == n - R && h.peer_hi != nullptr) {
      h.peer_hi[0] = out[1];
      h.peer_hi[1] = out[2];
      h.peer_hi[2] = out[3];
      __threadfence_system();
      asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(h.flag_hi), "l"(epoch + 1) : "memory");
    }
    if (c0 == 0 && h.epoch_out != nullptr) *h.epoch_out = epoch + 1;
  }
}

// Layout of the specialised stage kernel: shared differences (default) or the first form, which
// recomputes the neighbours' differences from shuffled cell values (same bits; A/B switch).
template <int EQ, int FLUX, int STAGE, bool WITH_MAX>
void launch_fast_layout(dim3 grid, int threads, cudaStream_t st, const FastParams &q) {
  if (g_fast_layout == 2)
    stage_warp_fast_share_kernel<EQ, FLUX, STAGE, WITH_MAX><<<grid, threads, 0, st>>>(q);
  else
    stage_warp_fast_kernel<EQ, FLUX, STAGE, WITH_MAX><<<grid, threads, 0, st>>>(q);
}

template <int EQ, int FLUX, int STAGE>
int launch_fast_stage(const StageParams &p, int batch, cudaStream_t st) {
  FastParams q{};
  q.uin = p.uin; q.u0 = p.u0; q.uout = p.uout; q.dt = p.dt; q.lf_speed = p.lf_speed;
  q.maxabs = p.maxabs; q.vel = p.vel; q.vel_l = p.vel_l; q.vel_r = p.vel_r; q.bc = p.bc; q.ld = p.ld;
  q.coef = p.invdx / FluxScale<EQ, FLUX>::value;
  q.eps9 = p.eps * (1.0 / 9.0);
  q.ca = p.ca; q.cb = p.cb; q.cc = p.cc;
  q.dt_stride = static_cast<int>(p.dt_stride);
  const FastGeometry geo = fast_geometry(p.bc.n, g_fast_wpc_max);
  q.chunks_per_row = geo.chunks_per_row;
  const int wpc = geo.wpc;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  unsigned gy, gz;
  if (!split_rows(batch, gy, gz)) return PSK_E_UNSUPPORTED;  // caller falls back
  const dim3 grid(gx, gy, gz);
  const int threads = wpc * 32;
  if (p.maxabs != nullptr)
    launch_fast_layout<EQ, FLUX, STAGE, true>(grid, threads, st, q);
  else
    launch_fast_layout<EQ, FLUX, STAGE, false>(grid, threads, st, q);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

template <int EQ, int FLUX>
int launch_fast(const StageParams &p, int batch, cudaStream_t st) {
  switch (p.stage) {
    case 0: return launch_fast_stage<EQ, FLUX, 0>(p, batch, st);
    case 1: return launch_fast_stage<EQ, FLUX, 1>(p, batch, st);
    case 2: return launch_fast_stage<EQ, FLUX, 2>(p, batch, st);
    case 4: return launch_fast_stage<EQ, FLUX, 4>(p, batch, st);
    default: return launch_fast_stage<EQ, FLUX, 3>(p, batch, st);
  }
}

// one chunk per warp
template <int EQ, int FLUX, int REC, bool STRICT, int R>
__global__ void __launch_bounds__(256)
stage_warp_kernel(const StageParams p, long long total_warps) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (warp_id >= total_warps) return;
  const ChunkRef c = chunk_ref<R>(p, warp_id, lane);
  if (!c.live) return;
  double own[R], u0v[R];
  chunk_load<R>(p, c, own, u0v);
  chunk_compute<EQ, FLUX, REC, STRICT, R>(p, c, lane, own, u0v);
}

// ---------------------------------------------------------------------------
// generic cell evaluation straight from global memory (zero padded, BC mapped): used for the
// ghost rows of a stage / RHS, and by the small parity kernels below.

template <int EQ, int FLUX, int REC, bool STRICT>
__device__ double rhs_at_cell(const StageParams &p, const double *__restrict__ urow, int row,
                              int i) {
  const int nx = p.bc.nx;
  double w[2 * kHalo + 3];  // w[i-4 .. i+4]
#pragma unroll
  for (int k = 0; k < 2 * kHalo + 3; ++k) w[k] = load_w(p.bc, urow, row, i - kHalo - 1 + k);
  // cells i-1, i, i+1 sit at window positions 3, 4, 5
  // ESWENO32: tau is zero at the two ends of the array (weno.py:292)
  const bool zm = (i - 1 <= 0) || (i - 1 >= nx - 1), zc = (i <= 0) || (i >= nx - 1), zp = (i + 1 <= 0) || (i + 1 >= nx - 1);
  Weno5Pair cm = reconstruct_cell<REC, STRICT>(w[1], w[2], w[3], w[4], w[5], p.eps, zm);
  Weno5Pair cc = reconstruct_cell<REC, STRICT>(w[2], w[3], w[4], w[5], w[6], p.eps, zc);
  Weno5Pair cp = reconstruct_cell<REC, STRICT>(w[3], w[4], w[5], w[6], w[7], p.eps, zp);
  double omm = 0.0, omc = 0.0, omp = 0.0;
  if (FLUX == PSK_FLUX_ESWENO) {
    omm = esweno32_cell<STRICT>(w[2], w[3], w[4], p.eps, zm).om0;
    omc = esweno32_cell<STRICT>(w[3], w[4], w[5], p.eps, zc).om0;
    omp = esweno32_cell<STRICT>(w[4], w[5], w[6], p.eps, zp).om0;
  }
  const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? p.lf_speed[row] : 0.0;
  double Flo = 0.0, Fhi = 0.0;  // jnp.pad(fnum, 1): the outermost faces carry zero flux
  if (i >= 1) {
    const int j = i - 1;
    double nu = (p.nu != nullptr) ? p.nu[j] : 1.0;
    double arj = (EQ != PSK_EQ_BURGERS) ? p.vel_r[j] : 0.0;
    double alp = (EQ != PSK_EQ_BURGERS) ? p.vel_l[j + 1] : 0.0;
    Flo = face_flux<EQ, FLUX, STRICT>(cm.ur, cc.ul, w[3], w[4], speed, nu, arj, alp);
    if (FLUX == PSK_FLUX_ESWENO) {
      const double gn = esweno_gnum<true>(omm, omc, w[3], w[4], p.delta);
      Flo = STRICT ? sadd(Flo, gn) : Flo + gn;
    }
  }
  if (i <= nx - 2) {
    const int j = i;
    double nu = (p.nu != nullptr) ? p.nu[j] : 1.0;
    double arj = (EQ != PSK_EQ_BURGERS) ? p.vel_r[j] : 0.0;
    double alp = (EQ != PSK_EQ_BURGERS) ? p.vel_l[j + 1] : 0.0;
    Fhi = face_flux<EQ, FLUX, STRICT>(cc.ur, cp.ul, w[4], w[5], speed, nu, arj, alp);
    if (FLUX == PSK_FLUX_ESWENO) {
      const double gn = esweno_gnum<true>(omc, omp, w[4], w[5], p.delta);
      Fhi = STRICT ? sadd(Fhi, gn) : Fhi + gn;
    }
  }
  const double vel = (EQ == PSK_EQ_ADVECTION) ? p.vel[i] : 0.0;
  return rhs_from_faces<EQ, STRICT>(Flo, Fhi, vel, p.dx, p.invdx);
}

// one thread per ghost cell: 2 g cells per row
template <int EQ, int FLUX, int REC, bool STRICT>
__global__ void ghost_rows_kernel(const StageParams p, int batch) {
  const int g = p.bc.g;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * 2 * g) return;
  const int row = idx / (2 * g);
  const int k = idx - row * 2 * g;
  if (p.active != nullptr && p.active[row] == 0) return;
  const int i = k < g ? k : p.bc.nx - 2 * g + k;
  const double *__restrict__ urow = p.uin + static_cast<int64_t>(row) * p.ld;
  const double L = rhs_at_cell<EQ, FLUX, REC, STRICT>(p, urow, row, i);
  const double dt = (p.stage != 0) ? p.dt[static_cast<int64_t>(row) * p.dt_stride] : 0.0;
  const int64_t off = static_cast<int64_t>(row) * p.ld + i;
  const double u0 = (p.stage >= 2) ? p.u0[off] : 0.0;
  // the stage combine uses the RAW stored value of uin, not the boundary-filled one
  p.uout[off] = stage_combine_p<STRICT>(p, u0, urow[i], dt, L);
}

// ---------------------------------------------------------------------------
// dispatch

// 0: warp kernel (default), 1: shared-memory tile kernel (kept for A/B measurements)
static int g_stage_variant = 0;
static int g_warp_block = 256;  // threads per CTA of the warp kernel (tuning)
static int g_use_fast = 1;  // specialised kernel for the hot configuration (stage_warp_fast_kernel)

template <int EQ, int FLUX, int REC, bool STRICT>
int launch_stage_warp(const StageParams &p, int batch, int ghost_rows, cudaStream_t st);

template <int EQ, int FLUX, int REC, bool STRICT>
int launch_stage(const StageParams &p, int batch, int ghost_rows, cudaStream_t st) {
  // the tile kernel has no ESWENO32 form
  if (g_stage_variant == 0 || REC == PSK_REC_ESWENO32)
    return launch_stage_warp<EQ, FLUX, REC, STRICT>(p, batch, ghost_rows, st);
  constexpr int R = 4;
  const int n = p.bc.n;
  int threads = (n + R - 1) / R;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  const int tile_cells = R * threads;
  StageParams q = p;
  q.tiles_per_row = (n + tile_cells - 1) / tile_cells;
  const bool aligned = (reinterpret_cast<uintptr_t>(p.uin + p.bc.g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p.uout + p.bc.g) % 16 == 0) &&
                       (p.u0 == nullptr || reinterpret_cast<uintptr_t>(p.u0 + p.bc.g) % 16 == 0) &&
                       (p.ld % 2 == 0);
  q.vec_ok = aligned ? 1 : 0;
  const size_t smem = sizeof(double) * (pad_index<R>(tile_cells + 2 * kHalo) + 1 +
                                        2 * (threads + 1) + 8 + 2);
  const long long blocks = static_cast<long long>(q.tiles_per_row) * batch;
  if (blocks > 2147483647LL) return PSK_E_INVALID;
  stage_tile_kernel<EQ, FLUX, REC, STRICT, R>
      <<<static_cast<unsigned>(blocks), threads, smem, st>>>(q);
  PSK_CUDA_OK(cudaGetLastError());
  if (ghost_rows) {
    const int total = batch * 2 * p.bc.g;
    ghost_rows_kernel<EQ, FLUX, REC, STRICT><<<(total + 127) / 128, 128, 0, st>>>(q, batch);
    PSK_CUDA_OK(cudaGetLastError());
  }
  return PSK_OK;
}

template <int EQ, int FLUX, int REC, bool STRICT>
int launch_stage_warp(const StageParams &p, int batch, int ghost_rows, cudaStream_t st) {
  constexpr int R = 4;
  StageParams q = p;
  q.tiles_per_row = (p.bc.n + 30 * R - 1) / (30 * R);  // warp chunks per row (2 halo lanes)
  const bool aligned = (reinterpret_cast<uintptr_t>(p.uin + p.bc.g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p.uout + p.bc.g) % 16 == 0) &&
                       (p.u0 == nullptr || reinterpret_cast<uintptr_t>(p.u0 + p.bc.g) % 16 == 0) &&
                       (p.ld % 2 == 0);
  q.vec_ok = aligned ? 1 : 0;
  if (g_use_fast && REC == PSK_REC_WENOJS53 && !STRICT && q.vec_ok && p.nu == nullptr && p.active == nullptr &&
      (batch <= 65535 || batch % 65535 == 0 || batch % 32768 == 0)) {
    int rc = PSK_E_UNSUPPORTED;
    unsigned ty, tz;
    if (split_rows(batch, ty, tz)) {
      rc = launch_fast<EQ, FLUX>(q, batch, st);
    } else {
      // rows beyond the grid.y limit: split the batch into equal slices of 32768 rows
      rc = PSK_OK;
      for (int b0 = 0; b0 < batch && rc == PSK_OK; b0 += 32768) {
        StageParams s2 = q;
        s2.uin = q.uin + static_cast<int64_t>(b0) * q.ld;
        s2.uout = q.uout + static_cast<int64_t>(b0) * q.ld;
        if (q.u0 != nullptr) s2.u0 = q.u0 + static_cast<int64_t>(b0) * q.ld;
        if (q.dt != nullptr) s2.dt = q.dt + static_cast<int64_t>(b0) * q.dt_stride;
        if (q.lf_speed != nullptr) s2.lf_speed = q.lf_speed + b0;
        if (q.maxabs != nullptr) s2.maxabs = q.maxabs + b0;
        if (q.bc.ghost != nullptr) s2.bc.ghost = q.bc.ghost + static_cast<int64_t>(b0) * q.bc.ghost_ld;
        rc = launch_fast<EQ, FLUX>(s2, 32768, st);
      }
    }
    if (rc == PSK_OK) {
      if (ghost_rows) {
        const int total = batch * 2 * p.bc.g;
        ghost_rows_kernel<EQ, FLUX, REC, STRICT><<<(total + 127) / 128, 128, 0, st>>>(q, batch);
        PSK_CUDA_OK(cudaGetLastError());
      }
      return PSK_OK;
    }
    if (rc != PSK_E_UNSUPPORTED) return rc;
  }
  const long long warps = static_cast<long long>(q.tiles_per_row) * batch;
  // small problems: fewer warps per CTA so that the chunks spread over more SMs
  int threads = g_warp_block;
  while (threads > 32 && warps * 32 / threads < 2 * kSMs) threads >>= 1;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  if (blocks > 2147483647LL) return PSK_E_INVALID;
  stage_warp_kernel<EQ, FLUX, REC, STRICT, R>
      <<<static_cast<unsigned>(blocks), threads, 0, st>>>(q, warps);
  PSK_CUDA_OK(cudaGetLastError());
  if (ghost_rows) {
    const int total = batch * 2 * p.bc.g;
    ghost_rows_kernel<EQ, FLUX, REC, STRICT><<<(total + 127) / 128, 128, 0, st>>>(q, batch);
    PSK_CUDA_OK(cudaGetLastError());
  }
  return PSK_OK;
}

template <int EQ, int FLUX, bool STRICT>
int dispatch_rec(int rec, const StageParams &p, int batch, int ghost_rows, cudaStream_t st) {
  switch (rec) {
    case PSK_REC_CONSTANT:
      return launch_stage<EQ, FLUX, PSK_REC_CONSTANT, STRICT>(p, batch, ghost_rows, st);
    case PSK_REC_WENOJS32:
      return launch_stage<EQ, FLUX, PSK_REC_WENOJS32, STRICT>(p, batch, ghost_rows, st);
    case PSK_REC_ESWENO32:
      return launch_stage<EQ, FLUX, PSK_REC_ESWENO32, STRICT>(p, batch, ghost_rows, st);
    default:
      return launch_stage<EQ, FLUX, PSK_REC_WENOJS53, STRICT>(p, batch, ghost_rows, st);
  }
}

template <bool STRICT>
int dispatch_scheme(const psk_desc *d, const StageParams &p, int ghost_rows, cudaStream_t st) {
  const int b = d->batch;
  if (d->equation == PSK_EQ_ADVECTION)
    return dispatch_rec<PSK_EQ_ADVECTION, PSK_FLUX_UPWIND, STRICT>(d->rec, p, b, ghost_rows, st);
  if (d->equation == PSK_EQ_CONTINUITY)
    return dispatch_rec<PSK_EQ_CONTINUITY, PSK_FLUX_UPWIND, STRICT>(d->rec, p, b, ghost_rows, st);
  switch (d->flux) {
    case PSK_FLUX_RUSANOV:
      return dispatch_rec<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV, STRICT>(d->rec, p, b, ghost_rows, st);
    case PSK_FLUX_LAX_FRIEDRICHS:
      return dispatch_rec<PSK_EQ_BURGERS, PSK_FLUX_LAX_FRIEDRICHS, STRICT>(d->rec, p, b,
                                                                           ghost_rows, st);
    case PSK_FLUX_UPWIND:
      return dispatch_rec<PSK_EQ_BURGERS, PSK_FLUX_UPWIND, STRICT>(d->rec, p, b, ghost_rows, st);
    case PSK_FLUX_ESWENO:  // check_desc: only with the ESWENO32 reconstruction
      return launch_stage<PSK_EQ_BURGERS, PSK_FLUX_ESWENO, PSK_REC_ESWENO32, STRICT>(p, b, ghost_rows, st);
    default:
      return dispatch_rec<PSK_EQ_BURGERS, PSK_FLUX_ENGQUIST_OSHER, STRICT>(d->rec, p, b,
                                                                           ghost_rows, st);
  }
}

static StageParams make_params(const psk_desc *d) {
  StageParams p{};
  p.nu = d->nu;
  p.vel = d->velocity;
  p.vel_l = d->vel_l;
  p.vel_r = d->vel_r;
  p.bc = make_bc_view(d);
  p.ld = d->ld;
  p.dx = d->dx;
  p.invdx = 1.0 / d->dx;
  p.eps = d->eps;
  p.delta = d->delta;
  return p;
}

// ---------------------------------------------------------------------------
// small parity kernels (one thread per output entry)

__global__ void apply_boundary_kernel(BcView b, const double *u, double *w, int64_t ld, int batch) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(batch) * b.nx) return;
  const int row = static_cast<int>(idx / b.nx);
  const int i = static_cast<int>(idx - static_cast<int64_t>(row) * b.nx);
  const double val = load_w(b, u + static_cast<int64_t>(row) * ld, row, i);
  if (w != u || i < b.g || i >= b.nx - b.g) w[static_cast<int64_t>(row) * ld + i] = val;
}

template <int REC, bool STRICT>
__global__ void reconstruct_kernel(const double *f, double *fl, double *fr, int nx, int64_t ld,
                                   int batch, double eps) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(batch) * nx) return;
  const int row = static_cast<int>(idx / nx);
  const int i = static_cast<int>(idx - static_cast<int64_t>(row) * nx);
  const double *r = f + static_cast<int64_t>(row) * ld;
  auto at = [&](int k) { return (k < 0 || k >= nx) ? 0.0 : r[k]; };
  Weno5Pair o = reconstruct_cell<REC, STRICT>(at(i - 2), at(i - 1), at(i), at(i + 1), at(i + 2), eps,
                                              i == 0 || i == nx - 1);
  fl[static_cast<int64_t>(row) * ld + i] = o.ul;
  fr[static_cast<int64_t>(row) * ld + i] = o.ur;
}

// F[k], k = 0..nx, from w (ghost cells already set): BcView with bc = NONE
template <int EQ, int FLUX, int REC, bool STRICT>
__global__ void numerical_flux_kernel(const StageParams p, double *F, int64_t ld_f, int batch) {
  const int nf = p.bc.nx + 1;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(batch) * nf) return;
  const int row = static_cast<int>(idx / nf);
  const int k = static_cast<int>(idx - static_cast<int64_t>(row) * nf);
  double val = 0.0;
  if (k >= 1 && k <= p.bc.nx - 1) {
    const int j = k - 1;
    const double *urow = p.uin + static_cast<int64_t>(row) * p.ld;
    double w[6];
#pragma unroll
    for (int m = 0; m < 6; ++m) w[m] = load_w(p.bc, urow, row, j - 2 + m);
    const bool za = (j == 0), zb = (j + 1 == p.bc.nx - 1);  // ESWENO32: tau = 0 at the array ends
    Weno5Pair a = reconstruct_cell<REC, STRICT>(w[0], w[1], w[2], w[3], w[4], p.eps, za);
    Weno5Pair b = reconstruct_cell<REC, STRICT>(w[1], w[2], w[3], w[4], w[5], p.eps, zb);
    const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? p.lf_speed[row] : 0.0;
    double nu = (p.nu != nullptr) ? p.nu[j] : 1.0;
    double arj = (EQ != PSK_EQ_BURGERS) ? p.vel_r[j] : 0.0;
    double alp = (EQ != PSK_EQ_BURGERS) ? p.vel_l[j + 1] : 0.0;
    val = face_flux<EQ, FLUX, STRICT>(a.ur, b.ul, w[2], w[3], speed, nu, arj, alp);
    if (FLUX == PSK_FLUX_ESWENO) {
      const double omj = esweno32_cell<STRICT>(w[1], w[2], w[3], p.eps, za).om0;
      const double omp = esweno32_cell<STRICT>(w[2], w[3], w[4], p.eps, zb).om0;
      const double gn = esweno_gnum<true>(omj, omp, w[2], w[3], p.delta);
      val = STRICT ? sadd(val, gn) : val + gn;
    }
  }
  F[static_cast<int64_t>(row) * ld_f + k] = val;
}

template <int EQ, int FLUX, bool STRICT>
int launch_flux_rec(int rec, const StageParams &p, double *F, int64_t ld_f, int batch,
                    cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(batch) * (p.bc.nx + 1);
  const unsigned blocks = static_cast<unsigned>((total + 127) / 128);
  switch (rec) {
    case PSK_REC_CONSTANT:
      numerical_flux_kernel<EQ, FLUX, PSK_REC_CONSTANT, STRICT><<<blocks, 128, 0, st>>>(p, F, ld_f, batch);
      break;
    case PSK_REC_WENOJS32:
      numerical_flux_kernel<EQ, FLUX, PSK_REC_WENOJS32, STRICT><<<blocks, 128, 0, st>>>(p, F, ld_f, batch);
      break;
    case PSK_REC_ESWENO32:
      numerical_flux_kernel<EQ, FLUX, PSK_REC_ESWENO32, STRICT><<<blocks, 128, 0, st>>>(p, F, ld_f, batch);
      break;
    default:
      numerical_flux_kernel<EQ, FLUX, PSK_REC_WENOJS53, STRICT><<<blocks, 128, 0, st>>>(p, F, ld_f, batch);
  }
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

template <bool STRICT>
int launch_flux(const psk_desc *d, const StageParams &p, double *F, int64_t ld_f, cudaStream_t st) {
  const int b = d->batch;
  if (d->equation == PSK_EQ_ADVECTION)
    return launch_flux_rec<PSK_EQ_ADVECTION, PSK_FLUX_UPWIND, STRICT>(d->rec, p, F, ld_f, b, st);
  if (d->equation == PSK_EQ_CONTINUITY)
    return launch_flux_rec<PSK_EQ_CONTINUITY, PSK_FLUX_UPWIND, STRICT>(d->rec, p, F, ld_f, b, st);
  switch (d->flux) {
    case PSK_FLUX_RUSANOV:
      return launch_flux_rec<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV, STRICT>(d->rec, p, F, ld_f, b, st);
    case PSK_FLUX_LAX_FRIEDRICHS:
      return launch_flux_rec<PSK_EQ_BURGERS, PSK_FLUX_LAX_FRIEDRICHS, STRICT>(d->rec, p, F, ld_f, b, st);
    case PSK_FLUX_UPWIND:
      return launch_flux_rec<PSK_EQ_BURGERS, PSK_FLUX_UPWIND, STRICT>(d->rec, p, F, ld_f, b, st);
    case PSK_FLUX_ESWENO:
      return launch_flux_rec<PSK_EQ_BURGERS, PSK_FLUX_ESWENO, STRICT>(PSK_REC_ESWENO32, p, F, ld_f, b, st);
    default:
      return launch_flux_rec<PSK_EQ_BURGERS, PSK_FLUX_ENGQUIST_OSHER, STRICT>(d->rec, p, F, ld_f, b, st);
  }
}

// max |w| per row; mode 0: all nx stored cells, 1: interior cells, 2: all nx cells after the
// boundary condition (the speed of the global Lax-Friedrichs flux, scalar.py:277)
__global__ void max_abs_kernel(BcView b, const double *u, int64_t ld, int mode,
                               unsigned long long *out, int chunks_per_row) {
  const int row = blockIdx.x / chunks_per_row;
  const int chunk = blockIdx.x - row * chunks_per_row;
  const double *urow = u + static_cast<int64_t>(row) * ld;
  const int lo = (mode == 1) ? b.g : 0;
  const int hi = (mode == 1) ? b.nx - b.g : b.nx;
  const int per = (hi - lo + chunks_per_row - 1) / chunks_per_row;
  const int a = lo + chunk * per;
  const int z = min(hi, a + per);
  unsigned long long m = 0ull;
  for (int i = a + threadIdx.x; i < z; i += blockDim.x) {
    const double val = (mode == 2) ? load_w(b, urow, row, i) : urow[i];
    const unsigned long long bits = abs_bits(val);
    m = bits > m ? bits : m;
  }
  __shared__ unsigned long long wm[32];
  m = warp_max_bits(m);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) wm[wid] = m;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    unsigned long long m2 = lane < nw ? wm[lane] : 0ull;
    m2 = warp_max_bits(m2);
    if (lane == 0) atomicMax(out + row, m2);
  }
}

static int launch_max_abs(const psk_desc *d, const double *u, int mode, double *out,
                          cudaStream_t st) {
  PSK_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * d->batch, st));
  const int nx = d->n + 2 * d->g;
  int chunks = 1;
  // enough CTAs to fill the machine when there are few rows
  while (static_cast<long long>(chunks) * d->batch < 2 * kSMs && nx / (chunks * 2) >= 2048) chunks *= 2;
  const int threads = nx >= 1024 ? 256 : 128;
  max_abs_kernel<<<static_cast<unsigned>(chunks) * d->batch, threads, 0, st>>>(
      make_bc_view(d), u, d->ld, mode, reinterpret_cast<unsigned long long *>(out), chunks);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

int launch_max_abs_public(const psk_desc *d, const double *u, int mode, double *out,
                          cudaStream_t st) {
  return launch_max_abs(d, u, mode, out, st);
}

// timestepping.py:139-150 on the device, one thread per row
__global__ void step_control_kernel(int batch, double theta, double cfl_scale, double tfinal,
                                    const double *maxabs, const double *t, double *t_next,
                                    double *dt, uint8_t *active, int *nonfinite) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= batch) return;
  const double tr = t[r];
  if (tr >= tfinal) {  // timestepping.py:133-134
    active[r] = 0;
    dt[r] = 0.0;
    t_next[r] = tr;
    return;
  }
  double step = __dmul_rn(theta, __ddiv_rn(cfl_scale, maxabs[r]));
  const double dt_min = __dadd_rn(tfinal, -tr);
  step = __dadd_rn(step < dt_min ? step : dt_min, 1.0e-15);  // timestepping.py:140-142
  if (!isfinite(step)) atomicOr(nonfinite, 1);                // timestepping.py:144-145
  active[r] = 1;
  dt[r] = step;
  t_next[r] = __dadd_rn(tr, step);
}

// ---- whole-step kernel (step_warp_fused_kernel): cells per lane and CTA shape, tuning switch
// psk_set_stage_variant(7000 + 10 R + shape): 66 = R 6, 32 x 16 (default), 61 = R 6, 32 x 12, 62 = R 6, 128 x 3, 60 = R 6, 256 x 2,
// 82 = R 8, 192 x 2;
// 7000 = off (three stage launches)
#ifndef PSK_STEP_MINB
#define PSK_STEP_MINB 16  // resident one-warp CTAs per SM the whole-step kernels are compiled for (16: within 128 registers)
#endif
static int g_step_variant = 66;  // R = 6, CTAs of ONE warp, 16 per SM (126 registers): 9.16e10 cell-updates/s on B200 (61: 12 per
                                 // SM at 130 registers, 9.03e10)
                                 // (62: CTAs of 128 threads, 3 per SM: 0.5 % slower -- a CTA waits for its slowest warp)

template <int R, int FLUX, int THREADS, int MINB, bool STAGES = false, int EQ = PSK_EQ_BURGERS, int BCK = 0, bool NU = false>
int launch_step_shape(const StepParams &q0, int n, int batch, bool with_max, cudaStream_t st) {
  StepParams q = q0;
  q.chunks_per_row = (n + StepGeometry<R>::kEmit - 1) / StepGeometry<R>::kEmit;
  int wpc = THREADS / 32;
  if (q.chunks_per_row < wpc) wpc = q.chunks_per_row;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  unsigned gy, gz;
  if (!split_rows(batch, gy, gz)) return PSK_E_UNSUPPORTED;
  const dim3 grid(gx, gy, gz);
  if (STAGES)
    step_warp_fused_kernel<R, FLUX, false, THREADS, MINB, STAGES, EQ, BCK, NU><<<grid, wpc * 32, 0, st>>>(q);
  else if (with_max)
    step_warp_fused_kernel<R, FLUX, true, THREADS, MINB, false, EQ, BCK, NU><<<grid, wpc * 32, 0, st>>>(q);
  else
    step_warp_fused_kernel<R, FLUX, false, THREADS, MINB, false, EQ, BCK, NU><<<grid, wpc * 32, 0, st>>>(q);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

// ---------------------------------------------------------------------------
// The whole SSPRK33 step with the GLOBAL Lax-Friedrichs flux (scalar.py:258-278) in one launch.  Every stage needs
// max |w| over the whole row of its input (scalar.py:277) -- a row-wide reduction between the stages, which is what
// keeps this flux out of step_warp_fused_kernel.  Here one thread-block CLUSTER owns a row (up to 8 CTAs of up to 12
// warps, one window per warp as in step_warp_fused_kernel) and the partial maxima travel through distributed shared
// memory:
//   * every warp reduces the stage input over its STORED cells (they partition the row) and sends the maximum to its
//     own slot in the shared memory of EVERY CTA of the cluster: st.async, whose bytes complete a transaction
//     mbarrier of the receiving CTA (one per stage, expecting 8 bytes from every window of the row);
//   * it then reconstructs the stage input -- most of the stage, none of which needs the speed -- and only waits on
//     its CTA's mbarrier right before the fluxes (LfSpeed below), where it reduces the row's slots.
// No cluster barrier and no memory fence between the stages (barrier.cluster.arrive.release costs a MEMBAR.GPU and
// an L1 invalidation each: measured 4.3-5.3e10 cell-updates/s with them, against 6.2e10 for four launches per step).
// One cluster barrier at the start publishes the mbarrier initialisation.  A CTA cannot exit before it has received
// everything addressed to it: its warps wait on all three mbarriers.
// Arithmetic: step_stage_rhs with the speed of the stage kernels, i.e. the bits of three psk_ssprk33_stage launches.
constexpr int kLfMaxWindows = 96;  // 8 CTAs x 12 warps

struct LfExchange {
  unsigned long long vals[3][kLfMaxWindows];  // bit patterns of the windows' max |stage input|
  unsigned long long mbar[3];
};

__device__ __forceinline__ unsigned smem_u32(const void *ptr) { return static_cast<unsigned>(__cvta_generic_to_shared(ptr)); }
__device__ __forceinline__ unsigned cluster_size_x() {
  unsigned v;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(v));
  return v;
}
// the address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ unsigned cluster_map(unsigned local, unsigned rank) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
  return remote;
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(mbar), "r"(parity) : "memory");
}

// max over the warp of 64-bit patterns in two 32-bit hardware reductions (redux.sync): the high words first, then
// the low words of the lanes that hold the largest high word -- 6 instructions instead of 5 shuffle rounds on pairs
__device__ __forceinline__ unsigned long long warp_max_bits_redux(unsigned long long v) {
  const unsigned hi = static_cast<unsigned>(v >> 32), lo = static_cast<unsigned>(v);
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  return (static_cast<unsigned long long>(mhi) << 32) | mlo;
}

struct LfSpeed {
  const LfExchange *x;
  int stage, windows, lane;
  // -2 max |w| of the row, once the maxima of all its windows have arrived in this CTA
  __device__ __forceinline__ double operator()() const {
    mbar_wait(smem_u32(&x->mbar[stage]), 0u);
    unsigned long long mx = 0ull;
#pragma unroll
    for (int k = 0; k < kLfMaxWindows / 32; ++k) {
      const int i = lane + 32 * k;
      const unsigned long long b = i < windows ? x->vals[stage][i] : 0ull;
      mx = b > mx ? b : mx;
    }
    return -2.0 * __longlong_as_double(static_cast<long long>(warp_max_bits_redux(mx)));
  }
};

// `extra`: what else enters the maximum (Dirichlet rows, first window: |boundary data| of this stage, loaded at the
// start of the kernel -- a load here would sit on the critical path of every warp of the row)
template <int R>
__device__ __forceinline__ void lf_send_stage_max(LfExchange *x, int lane, int chunk, int stage, unsigned long long extra,
                                                  const bool (&st)[R], const double (&v)[R]) {
  unsigned long long mx = extra;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (st[r]) {
      const unsigned long long b = abs_bits(v[r]);
      mx = b > mx ? b : mx;
    }
  mx = warp_max_bits_redux(mx);
  const unsigned nc = cluster_size_x();
  if (nc == 1) {  // a row of one CTA: a plain store and an arrival (the mbarrier then counts the windows, not bytes)
    if (lane == 0) {
      x->vals[stage][chunk] = mx;
      asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&x->mbar[stage])) : "memory");
    }
  } else if (static_cast<unsigned>(lane) < nc) {  // lane r: to CTA r of the cluster
    const unsigned dst = cluster_map(smem_u32(&x->vals[stage][chunk]), static_cast<unsigned>(lane));
    const unsigned bar = cluster_map(smem_u32(&x->mbar[stage]), static_cast<unsigned>(lane));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(mx), "r"(bar)
                 : "memory");
  }
}

// step_fill_ghosts from the warp's copy of the row's 2 g <= 32 boundary data of one stage
template <int R>
__device__ __forceinline__ void lf_fill_ghosts(const StepParams &p, const double (&gh)[32], int c0, double (&a)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = c0 + r;
    if (c < 0) a[r] = (c >= -p.g) ? gh[c + p.g] : 0.0;
    if (c >= p.n) a[r] = (c < p.n + p.g) ? gh[p.g + c - p.n] : 0.0;
  }
}

// STAGES: the stage values k1, k2 are stored too (p.k1_out, p.k2_out: what the reverse sweep recomputes from a
// checkpointed state); with p.uout == nullptr the third stage is skipped -- nothing is sent for it and nobody waits
template <int R, bool WITH_MAX, int BCK, int THREADS, int MINB, bool NU, bool STAGES = false>
__global__ void __launch_bounds__(THREADS, MINB)
step_lf_cluster_kernel(const StepParams p) {
  using Geo = StepGeometry<R>;
  constexpr int kLF = PSK_FLUX_LAX_FRIEDRICHS, kB = PSK_EQ_BURGERS;
  constexpr bool DIRICHLET = BCK == 1;
  __shared__ LfExchange xch;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool live = chunk < p.chunks_per_row;
  const int row = blockIdx.y + blockIdx.z * gridDim.y;
  const int n = p.n;
  const int c0 = chunk * Geo::kEmit - Geo::kSkip + R * lane;
  const bool inside = (c0 >= 0) && (c0 + R <= n);
  const int64_t base = static_cast<int64_t>(row) * p.ld + p.g;
  const int wc0 = R * lane;
  bool st[R];
#pragma unroll
  for (int r = 0; r < R; ++r)
    st[r] = live && (wc0 + r >= Geo::kSkip) && (wc0 + r < Geo::kWindow - Geo::kSkip) && (c0 + r >= 0) && (c0 + r < n);

  const bool skip = p.active != nullptr && p.active[row] == 0;  // finished row (the whole cluster agrees)
  if (threadIdx.x == 0 && !skip) {  // one transaction barrier per stage: 8 bytes from every window of the row
    const bool solo = cluster_size_x() == 1;  // (a row of one CTA: one arrival per window instead, see lf_send_stage_max)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const unsigned bar = smem_u32(&xch.mbar[s]);
      if (solo) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(static_cast<unsigned>(p.chunks_per_row)) : "memory");
      } else {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                     "r"(8u * static_cast<unsigned>(p.chunks_per_row))
                     : "memory");
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  // Dirichlet rows: w = apply_boundary(u) holds the boundary data of the stage in its ghost cells (scalar.py:277)
  // Every warp whose window reaches beyond the row keeps the data of the three stage times in shared memory: the
  // fills in front of stages 2 and 3 must not wait for global loads -- the WHOLE row waits for its slowest warp.
  __shared__ double gdata[DIRICHLET ? 12 : 1][3][32];
  double(&gmine)[3][32] = gdata[DIRICHLET ? (threadIdx.x >> 5) : 0];
  unsigned long long gbits[3] = {0ull, 0ull, 0ull};
  const bool edge = live && ((chunk == 0) || ((chunk + 1) * Geo::kEmit + Geo::kSkip > n));
  if (DIRICHLET && edge && lane < 2 * p.g && !skip) {
    const double *gh = p.ghost3 + static_cast<int64_t>(row) * p.ghost_ld + lane;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const double v = gh[s * p.ghost_block];
      gmine[s][lane] = v;
      if (chunk == 0) gbits[s] = abs_bits(v);
    }
  }
  double u0[R];
#pragma unroll
  for (int r = 0; r < R; ++r) u0[r] = 0.0;
  if (live) {
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
        if (DIRICHLET) {
          u0[r] = (c >= 0 && c < n) ? p.u[base + c] : 0.0;
        } else {
          c %= n;
          if (c < 0) c += n;
          u0[r] = p.u[base + c];
        }
      }
    }
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // every CTA's mbarriers are ready
  if (skip) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st[r]) p.uout[base + c0 + r] = u0[r];
    return;
  }
  if (!live) return;  // a warp beyond the last window of the row: sends nothing, nobody waits for it
  const double cdt = p.coef * p.dt[static_cast<int64_t>(row) * p.dt_stride];
  const bool fill = DIRICHLET && !inside;
  const int windows = p.chunks_per_row;
  double nuf[R + 1];
  if constexpr (NU) step_load_nu<R>(p, c0, nuf);

  double a[R], dF[R];
  // (the maximum goes out BEFORE the ghost cells of this warp are filled: stored cells only enter it, and the loads
  // of the fill stay off the critical path of the row)
  if (DIRICHLET) __syncwarp();
  lf_send_stage_max<R>(&xch, lane, chunk, 0, gbits[0], st, u0);
  if (fill) lf_fill_ghosts<R>(p, gmine[0], c0, u0);
  step_stage_rhs<R, kLF, kB, LfSpeed, NU>(u0, p.eps9, dF, nullptr, LfSpeed{&xch, 0, windows, lane}, nuf);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(cdt, dF[r], u0[r]);  // k1
  auto store_cells = [&](double *dst, const double(&v)[R]) {
    if (inside) {
#pragma unroll
      for (int r = 0; r < R; r += 2)
        if (st[r]) *reinterpret_cast<double2 *>(dst + base + c0 + r) = make_double2(v[r], v[r + 1]);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (st[r]) dst[base + c0 + r] = v[r];
    }
  };
  if (STAGES) store_cells(p.k1_out, a);
  lf_send_stage_max<R>(&xch, lane, chunk, 1, gbits[1], st, a);
  if (fill) lf_fill_ghosts<R>(p, gmine[1], c0, a);
  step_stage_rhs<R, kLF, kB, LfSpeed, NU>(a, p.eps9, dF, nullptr, LfSpeed{&xch, 1, windows, lane}, nuf);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(0.25, fma(cdt, dF[r], a[r]), 0.75 * u0[r]);  // k2
  if (STAGES) {
    store_cells(p.k2_out, a);
    if (p.uout == nullptr) return;  // (every message addressed to this CTA arrived before its warps left stage 2)
  }
  lf_send_stage_max<R>(&xch, lane, chunk, 2, gbits[2], st, a);
  if (fill) lf_fill_ghosts<R>(p, gmine[2], c0, a);
  step_stage_rhs<R, kLF, kB, LfSpeed, NU>(a, p.eps9, dF, nullptr, LfSpeed{&xch, 2, windows, lane}, nuf);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(2.0 / 3.0, fma(cdt, dF[r], a[r]), (1.0 / 3.0) * u0[r]);  // u'

  store_cells(p.uout, a);
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

// cluster shape of a row: CTAs of g_lf_wpc windows (warps) each -- small CTAs, several rows resident per SM, so that
// the rows' load / compute / barrier phases overlap --, more windows per CTA only where a row would need more than
// the 8 CTAs of a portable cluster
static int g_lf_wpc = 4;

template <bool WITH_MAX, int BCK, bool NU = false, bool STAGES = false>
int launch_step_lf(const StepParams &q0, int n, int batch, cudaStream_t st) {
  constexpr int R = 6, kMaxWarps = 12;
  StepParams q = q0;
  q.chunks_per_row = (n + StepGeometry<R>::kEmit - 1) / StepGeometry<R>::kEmit;
  if (q.g > 16) return PSK_E_UNSUPPORTED;
  unsigned gy, gz;
  if (!split_rows(batch, gy, gz)) return PSK_E_UNSUPPORTED;
  // try 0: small CTAs in a cluster of up to 16 (beyond the portable 8: rows of 5505 .. 11 008 cells keep the 4-warp
  // CTAs that way); try 1: the portable cluster of at most 8 CTAs, as many windows per CTA as that takes
  static bool big_clusters_ok = true;
  for (int attempt = big_clusters_ok ? 0 : 1; attempt < 2; ++attempt) {
    const int cmax = attempt == 0 ? 16 : 8;
    int wpc = g_lf_wpc;
    if (wpc * cmax < q.chunks_per_row) wpc = (q.chunks_per_row + cmax - 1) / cmax;
    if (wpc > kMaxWarps) return PSK_E_UNSUPPORTED;  // rows beyond 8 x 12 x 172 = 16512 cells
    if (wpc > q.chunks_per_row) wpc = q.chunks_per_row;
    const int cx = (q.chunks_per_row + wpc - 1) / wpc;
    if (attempt == 0 && (cx <= 8 || wpc > 4)) continue;  // nothing gained over the portable shape
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(cx), gy, gz);
    cfg.blockDim = dim3(static_cast<unsigned>(wpc * 32));
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(cx);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // (CTAs of up to 4 windows: 16 warps per SM within 128 registers; longer rows: one large CTA per SM)
    cudaError_t err;
    if (wpc <= 4) {
      auto kernel = step_lf_cluster_kernel<R, WITH_MAX, BCK, 128, 4, NU, STAGES>;
      if (cx > 8) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (err == cudaSuccess) err = cudaLaunchKernelEx(&cfg, kernel, q);
      } else {
        err = cudaLaunchKernelEx(&cfg, kernel, q);
      }
    } else {
      err = cudaLaunchKernelEx(&cfg, step_lf_cluster_kernel<R, WITH_MAX, BCK, 384, 1, NU, STAGES>, q);
    }
    if (err == cudaSuccess) return PSK_OK;
    if (attempt == 0) {  // this device does not place clusters of that size: never ask again
      (void)cudaGetLastError();
      big_clusters_ok = false;
      continue;
    }
    PSK_CUDA_OK(err);
  }
  return PSK_OK;
}

// ---------------------------------------------------------------------------
// The whole-step kernel FUSED with the ghost-cell exchange of a slab-decomposed grid (one row per
// GPU, boundary kind NONE, 9 ghost cells per side: the three stages of a step reach 9 cells beyond
// the slab): ONE launch per step and nothing else on the exchange path.
//   * the two warps whose windows reach into ghost cells (chunk 0 and the last chunk; the host makes
//     sure that the last chunk holds at least 10 cells, so it is the only one on its side) spin on
//     the LOCAL epoch flags until the neighbours' edge cells of the current state have arrived --
//     every other warp of the grid starts at once, the NVLink round trip hides behind the interior;
//   * ghost cells are read with ld.volatile;
//   * the lanes that store the slab's first / last 9 cells also store them into the left / right
//     neighbour's ghost slots of the array the new state lives in, then one lane of the warp raises
//     that neighbour's flag to epoch + 1 (stores, __threadfence_system, __syncwarp, st.release.sys).
// No write-after-read hazard: the neighbour last read those ghost slots one step ago, in the very warp
// whose push of that step this warp has just waited for.  Arithmetic: step_stage_rhs, i.e. the bits of
// step_warp_fused_kernel and of three stage launches.
// the cells `mask` selects among a lane's six go to dst[0..5]; then (whole warp) the neighbour's flag is raised
__device__ __noinline__ void step_push_edge(double *dst, long long *flag, long long next_epoch, unsigned mask, int lane,
                                            double a0, double a1, double a2, double a3, double a4, double a5) {
  const double a[6] = {a0, a1, a2, a3, a4, a5};
#pragma unroll
  for (int r = 0; r < 6; ++r)
    if (mask & (1u << r)) dst[r] = a[r];
  if (mask != 0u) __threadfence_system();
  __syncwarp(0xffffffffu);
  if (lane == 0) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(flag), "l"(next_epoch) : "memory");
  }
}

template <int R, int FLUX, bool WITH_MAX, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
step_warp_fused_p2p_kernel(const StepParams p, const HaloLink h) {
  using Geo = StepGeometry<R>;
  static_assert(R == 6, "step_push_edge takes six cells");
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chunk >= p.chunks_per_row) return;
  const int n = p.n, g = p.g;
  // p.shift: the chunk grid starts `shift` cells left of the slab, so that BOTH end chunks hold >= 10 cells
  const int c0 = chunk * Geo::kEmit - Geo::kSkip - p.shift + R * lane;
  const bool inside = (c0 >= 0) && (c0 + R <= n);
  const int64_t base = g;
  const bool edge_lo = (chunk == 0), edge_hi = (chunk == p.chunks_per_row - 1);
  const long long epoch = h.wait_epoch;
  if (edge_lo && h.wait_lo != nullptr) halo_spin(h.wait_lo, epoch, h.timeout_ns, h.timed_out);
  if (edge_hi && h.wait_hi != nullptr) halo_spin(h.wait_hi, epoch, h.timeout_ns, h.timed_out);
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
    const volatile double *urow = p.u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = c0 + r;
      u0[r] = (c >= -g && c < n + g) ? urow[base + c] : 0.0;
    }
  }
  const double cdt = p.coef * p.dt[0];
  double a[R], dF[R];
  step_stage_rhs<R, FLUX>(u0, p.eps9, dF);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(cdt, dF[r], u0[r]);
  step_stage_rhs<R, FLUX>(a, p.eps9, dF);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(0.25, fma(cdt, dF[r], a[r]), 0.75 * u0[r]);
  step_stage_rhs<R, FLUX>(a, p.eps9, dF);
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = fma(2.0 / 3.0, fma(cdt, dF[r], a[r]), (1.0 / 3.0) * u0[r]);
  step_store<R>(p.uout + base + c0, inside, st, a);
  // ---- the slab's first / last 9 cells go to the neighbours' ghost slots, then their flags
  if (edge_lo && h.peer_lo != nullptr) {
    unsigned mask = 0u;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st[r] && c0 + r < 9) mask |= 1u << r;
    step_push_edge(h.peer_lo + c0, h.flag_lo, epoch + 1, mask, lane, a[0], a[1], a[2], a[3], a[4], a[5]);
  }
  if (edge_hi && h.peer_hi != nullptr) {
    unsigned mask = 0u;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st[r] && c0 + r >= n - 9) mask |= 1u << r;
    step_push_edge(h.peer_hi + (c0 - (n - 9)), h.flag_hi, epoch + 1, mask, lane, a[0], a[1], a[2], a[3], a[4], a[5]);
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
    if (lane == 0) atomicMax(p.maxabs, mx);
  }
}

template <int FLUX>
int launch_step_p2p(const StepParams &q, const HaloLink &h, bool with_max, cudaStream_t st) {
#ifndef PSK_P2P_MINB
#define PSK_P2P_MINB 4  // CTAs of four warps per SM: 4 = within 128 registers (16 warps per SM), 3 = up to 168
#endif
  constexpr int kThreads = 128, kMinB = PSK_P2P_MINB;
  int wpc = kThreads / 32;
  if (q.chunks_per_row < wpc) wpc = q.chunks_per_row;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  if (with_max)
    step_warp_fused_p2p_kernel<6, FLUX, true, kThreads, kMinB><<<gx, wpc * 32, 0, st>>>(q, h);
  else
    step_warp_fused_p2p_kernel<6, FLUX, false, kThreads, kMinB><<<gx, wpc * 32, 0, st>>>(q, h);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

int launch_step_fused(const psk_desc *d, const double *u, double *uout, const double *dt, int64_t dt_stride,
                      const uint8_t *active, double *maxabs, cudaStream_t st, double *k1_out = nullptr,
                      double *k2_out = nullptr, const double *ghost3 = nullptr) {
  StepParams q{};
  q.u = u; q.uout = uout; q.dt = dt; q.active = active;
  q.ghost3 = ghost3;
  q.ghost_ld = d->ghost_ld;
  q.ghost_block = d->ghost_ld != 0 ? static_cast<int64_t>(d->batch) * d->ghost_ld : 2 * d->g;
  q.vel = d->velocity; q.vel_l = d->vel_l; q.vel_r = d->vel_r;
  q.nu = d->nu;
  q.k1_out = k1_out; q.k2_out = k2_out;
  q.maxabs = reinterpret_cast<unsigned long long *>(maxabs);
  q.ld = d->ld;
  q.coef = (1.0 / d->dx) / (d->equation != PSK_EQ_BURGERS ? 1.0
                           : ((d->flux == PSK_FLUX_RUSANOV || d->flux == PSK_FLUX_LAX_FRIEDRICHS) ? 4.0 : 2.0));  // FluxScale
  q.eps9 = d->eps * (1.0 / 9.0);
  q.dt_stride = static_cast<int>(dt_stride);
  q.n = d->n;
  q.g = d->g;
  q.bc_none = d->bc == PSK_BC_NONE ? 1 : 0;
  const bool mx = maxabs != nullptr;
  int batch = d->batch;
  // rows beyond the grid limits with no even split over grid.y x grid.z: equal slices of 32768 rows
  unsigned ty, tz;
  if (!split_rows(batch, ty, tz)) {
    if (batch % 32768 != 0) return PSK_E_UNSUPPORTED;
    for (int b0 = 0; b0 < batch; b0 += 32768) {
      psk_desc d2 = *d;
      d2.batch = 32768;
      const int64_t o = static_cast<int64_t>(b0) * d->ld;
      const int rc = launch_step_fused(&d2, u + o, uout != nullptr ? uout + o : nullptr,
                                       dt + static_cast<int64_t>(b0) * dt_stride, dt_stride,
                                       active != nullptr ? active + b0 : nullptr,
                                       maxabs != nullptr ? maxabs + b0 : nullptr, st,
                                       k1_out != nullptr ? k1_out + o : nullptr, k2_out != nullptr ? k2_out + o : nullptr,
                                       (ghost3 != nullptr && d->ghost_ld != 0) ? ghost3 + static_cast<int64_t>(b0) * d->ghost_ld
                                                                               : ghost3);
      if (rc != PSK_OK) return rc;
    }
    return PSK_OK;
  }
  if (d->flux == PSK_FLUX_LAX_FRIEDRICHS) {  // one cluster per row (the entry points admit periodic / Dirichlet rows)
    if (k1_out != nullptr) {  // stage values for the reverse sweep (Dirichlet rows: psk_ssprk33_step_bc)
      if (d->bc != PSK_BC_DIRICHLET) return PSK_E_UNSUPPORTED;
      return d->nu != nullptr ? launch_step_lf<false, 1, true, true>(q, d->n, batch, st)
                              : launch_step_lf<false, 1, false, true>(q, d->n, batch, st);
    }
    if (d->bc == PSK_BC_DIRICHLET && d->nu != nullptr)  // alpha != 1 (the reference's burgers-adjoint defaults)
      return mx ? launch_step_lf<true, 1, true>(q, d->n, batch, st) : launch_step_lf<false, 1, true>(q, d->n, batch, st);
    if (d->bc == PSK_BC_DIRICHLET)
      return mx ? launch_step_lf<true, 1>(q, d->n, batch, st) : launch_step_lf<false, 1>(q, d->n, batch, st);
    return mx ? launch_step_lf<true, 0>(q, d->n, batch, st) : launch_step_lf<false, 0>(q, d->n, batch, st);
  }
  // rows with boundary data: CTAs of ONE warp, 12 per SM -- the two edge warps of a row are slower than the rest, and
  // in a CTA of four warps the other three wait for them with their registers held (ncu: 10.05 warps resident per SM
  // instead of 11.03 on periodic rows; Rusanov Dirichlet 8.6e10 -> 8.8e10)
  if (d->bc == PSK_BC_DIRICHLET || d->bc == PSK_BC_NEUMANN) {
    constexpr int kB = PSK_EQ_BURGERS, kUp = PSK_FLUX_UPWIND, kEO = PSK_FLUX_ENGQUIST_OSHER, kRus = PSK_FLUX_RUSANOV;
    const bool neumann = d->bc == PSK_BC_NEUMANN;
    const bool mxs = (k1_out != nullptr) ? false : mx;
    // (k1_out given: the stage values are stored too, for the reverse sweep)
#define PSK_STEP_BC(FL, EQ)                                                                                      \
  do {                                                                                                           \
    if (k1_out != nullptr)                                                                                       \
      return neumann ? launch_step_shape<6, FL, 32, PSK_STEP_MINB, true, EQ, 2>(q, d->n, batch, false, st)                  \
                     : launch_step_shape<6, FL, 32, PSK_STEP_MINB, true, EQ, 1>(q, d->n, batch, false, st);                 \
    return neumann ? launch_step_shape<6, FL, 32, PSK_STEP_MINB, false, EQ, 2>(q, d->n, batch, mxs, st)                     \
                   : launch_step_shape<6, FL, 32, PSK_STEP_MINB, false, EQ, 1>(q, d->n, batch, mxs, st);                    \
  } while (0)
    if (d->equation == PSK_EQ_ADVECTION) PSK_STEP_BC(kUp, PSK_EQ_ADVECTION);
    if (d->equation == PSK_EQ_CONTINUITY) PSK_STEP_BC(kUp, PSK_EQ_CONTINUITY);
    if (d->flux == PSK_FLUX_UPWIND) PSK_STEP_BC(kUp, kB);
    if (d->flux == PSK_FLUX_ENGQUIST_OSHER) PSK_STEP_BC(kEO, kB);
    if (d->nu != nullptr) {  // Rusanov with alpha != 1: nu of every face
      if (k1_out != nullptr)
        return neumann ? launch_step_shape<6, kRus, 32, PSK_STEP_MINB, true, kB, 2, true>(q, d->n, batch, false, st)
                       : launch_step_shape<6, kRus, 32, PSK_STEP_MINB, true, kB, 1, true>(q, d->n, batch, false, st);
      return neumann ? launch_step_shape<6, kRus, 32, PSK_STEP_MINB, false, kB, 2, true>(q, d->n, batch, mxs, st)
                     : launch_step_shape<6, kRus, 32, PSK_STEP_MINB, false, kB, 1, true>(q, d->n, batch, mxs, st);
    }
    PSK_STEP_BC(kRus, kB);
#undef PSK_STEP_BC
  }
  if (d->equation != PSK_EQ_BURGERS) {  // periodic rows, upwind flux with the velocity's reconstruction
    if (k1_out != nullptr) return PSK_E_UNSUPPORTED;
    if (d->equation == PSK_EQ_ADVECTION)
      return launch_step_shape<6, PSK_FLUX_UPWIND, 32, PSK_STEP_MINB, false, PSK_EQ_ADVECTION, false>(q, d->n, batch, mx, st);
    return launch_step_shape<6, PSK_FLUX_UPWIND, 32, PSK_STEP_MINB, false, PSK_EQ_CONTINUITY, false>(q, d->n, batch, mx, st);
  }
  if (k1_out != nullptr)  // stage values wanted (reverse sweep): Rusanov, default shape
    return launch_step_shape<6, PSK_FLUX_RUSANOV, 32, PSK_STEP_MINB, true>(q, d->n, batch, false, st);
  // the other Burgers fluxes: default shape only
  if (d->flux == PSK_FLUX_UPWIND) return launch_step_shape<6, PSK_FLUX_UPWIND, 32, PSK_STEP_MINB>(q, d->n, batch, mx, st);
  if (d->flux == PSK_FLUX_ENGQUIST_OSHER)
    return launch_step_shape<6, PSK_FLUX_ENGQUIST_OSHER, 32, PSK_STEP_MINB>(q, d->n, batch, mx, st);
  constexpr int kRus = PSK_FLUX_RUSANOV;
  switch (g_step_variant) {
    case 60: return launch_step_shape<6, kRus, 256, 2>(q, d->n, batch, mx, st);
    case 62: return launch_step_shape<6, kRus, 128, 3>(q, d->n, batch, mx, st);
    case 61: return launch_step_shape<6, kRus, 32, 12>(q, d->n, batch, mx, st);
    case 64: return launch_step_shape<6, kRus, 64, 6>(q, d->n, batch, mx, st);
    case 66: return launch_step_shape<6, kRus, 32, 16>(q, d->n, batch, mx, st);  // 128 registers, 16 warps per SM
    case 68: return launch_step_shape<6, kRus, 32, 20>(q, d->n, batch, mx, st);  // 96 registers, 21 warps per SM
    case 69: return launch_step_shape<6, kRus, 32, 24>(q, d->n, batch, mx, st);
    case 82: return launch_step_shape<8, kRus, 192, 2>(q, d->n, batch, mx, st);
    default: return PSK_E_UNSUPPORTED;
  }
}

}  // namespace psk

// ===========================================================================
// C ABI

using namespace psk;

extern "C" {

int psk_version(void) { return PSK_VERSION; }

/* tuning / A-B switch, not part of the reference-facing surface: 0 = warp-shuffle stage
 * kernel (default), 1 = shared-memory tile kernel */
int psk_set_stage_variant(int variant) {
  if (variant >= 8000) {  // 8000 + windows per CTA of the Lax-Friedrichs cluster kernel (1..12)
    if (variant - 8000 < 1 || variant - 8000 > 12) return PSK_E_INVALID;
    g_lf_wpc = variant - 8000;
    return PSK_OK;
  }
  if (variant >= 7000) {  // whole-step kernel (psk_ssprk33_step): 7066 (default) / 7061 / 7062 / 7064 / 7060 / 7082 = cells per lane and CTA shape, 7000 = off
    const int v = variant - 7000;
    if (v != 0 && v != 60 && v != 61 && v != 62 && v != 64 && v != 66 && v != 68 && v != 69 && v != 82) return PSK_E_INVALID;
    g_step_variant = v;
    return PSK_OK;
  }
  if (variant >= 5000) {  // 5000 + layout of the specialised stage kernel (0 or 2)
    const int layout = variant - 5000;
    if (layout != 0 && layout != 2) return PSK_E_INVALID;
    g_fast_layout = layout;
    return PSK_OK;
  }
  if (variant >= 4000) {  // 4000 + max warps per CTA of the specialised kernel (1..8)
    if (variant - 4000 < 1 || variant - 4000 > 8) return PSK_E_INVALID;
    g_fast_wpc_max = variant - 4000;
    return PSK_OK;
  }
  if (variant >= 3000) {  // 3000 / 3001: specialised fast kernel off / on
    g_use_fast = (variant - 3000) != 0;
    return PSK_OK;
  }
  if (variant >= 100) {  // 100 + threads per CTA of the warp kernel (32..256)
    const int t = variant - 100;
    if (t < 32 || t > 256 || (t % 32) != 0) return PSK_E_INVALID;
    g_warp_block = t;
    return PSK_OK;
  }
  if (variant < 0 || variant > 1) return PSK_E_INVALID;
  g_stage_variant = variant;
  return PSK_OK;
}

const char *psk_status_string(int status) {
  switch (status) {
    case PSK_OK: return "ok";
    case PSK_E_INVALID: return "invalid argument";
    case PSK_E_UNSUPPORTED: return "outside the supported hot path";
    case PSK_E_CUDA: return "CUDA runtime error";
    case PSK_E_NONFINITE: return "time step is not finite";
    default: return "unknown status";
  }
}

int psk_last_cuda_error(void) { return g_last_cuda_error; }

int psk_apply_boundary(const psk_desc *d, const double *u, double *w, psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (u == nullptr || w == nullptr) return PSK_E_INVALID;
  const BcView b = make_bc_view(d);
  const int64_t total = static_cast<int64_t>(d->batch) * b.nx;
  apply_boundary_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(b, u, w, d->ld, d->batch);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

int psk_reconstruct(const psk_desc *d, const double *f, double *fl, double *fr,
                    psk_stream_t stream) {
  if (d == nullptr || f == nullptr || fl == nullptr || fr == nullptr) return PSK_E_INVALID;
  if (d->n <= 0 || d->batch <= 0 || d->g < 0 || d->ld < d->n + 2 * d->g) return PSK_E_INVALID;
  if (d->rec < PSK_REC_CONSTANT || d->rec > PSK_REC_ESWENO32) return PSK_E_UNSUPPORTED;
  const int nx = d->n + 2 * d->g;
  const int64_t total = static_cast<int64_t>(d->batch) * nx;
  const unsigned blocks = static_cast<unsigned>((total + 127) / 128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool strict = d->math == PSK_MATH_STRICT;
#define PSK_REC_LAUNCH(REC)                                                                  \
  if (strict)                                                                                \
    reconstruct_kernel<REC, true><<<blocks, 128, 0, st>>>(f, fl, fr, nx, d->ld, d->batch, d->eps); \
  else                                                                                       \
    reconstruct_kernel<REC, false><<<blocks, 128, 0, st>>>(f, fl, fr, nx, d->ld, d->batch, d->eps)
  switch (d->rec) {
    case PSK_REC_CONSTANT: PSK_REC_LAUNCH(PSK_REC_CONSTANT); break;
    case PSK_REC_WENOJS32: PSK_REC_LAUNCH(PSK_REC_WENOJS32); break;
    case PSK_REC_ESWENO32: PSK_REC_LAUNCH(PSK_REC_ESWENO32); break;
    default: PSK_REC_LAUNCH(PSK_REC_WENOJS53);
  }
#undef PSK_REC_LAUNCH
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

int psk_max_abs(const psk_desc *d, const double *u, int interior_only, double *out,
                psk_stream_t stream) {
  if (d == nullptr || u == nullptr || out == nullptr) return PSK_E_INVALID;
  if (d->n <= 0 || d->batch <= 0 || d->g < 0 || d->ld < d->n + 2 * d->g) return PSK_E_INVALID;
  if (interior_only == 2) {
    int rc = check_desc(d);
    if (rc != PSK_OK) return rc;
  }
  return launch_max_abs(d, u, interior_only, out, static_cast<cudaStream_t>(stream));
}

// global Lax-Friedrichs speed max |w| over all nx cells after the BC (scalar.py:277)
static int lf_speed_pass(const psk_desc *d, const double *u, int bc_applied, double *lf_work,
                         cudaStream_t st) {
  if (d->equation != PSK_EQ_BURGERS || d->flux != PSK_FLUX_LAX_FRIEDRICHS) return PSK_OK;
  if (lf_work == nullptr) return PSK_E_INVALID;
  return launch_max_abs(d, u, bc_applied ? 0 : 2, lf_work, st);
}

int psk_numerical_flux(const psk_desc *d, const double *w, double *flux, int64_t ld_f,
                       double *lf_work, psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (w == nullptr || flux == nullptr || ld_f < d->n + 2 * d->g + 1) return PSK_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = lf_speed_pass(d, w, 1, lf_work, st);
  if (rc != PSK_OK) return rc;
  StageParams p = make_params(d);
  p.bc.bc = PSK_BC_NONE;  // the caller has applied the boundary condition already
  p.uin = w;
  p.lf_speed = lf_work;
  return d->math == PSK_MATH_STRICT ? launch_flux<true>(d, p, flux, ld_f, st)
                                    : launch_flux<false>(d, p, flux, ld_f, st);
}

int psk_ssprk33_stage(const psk_desc *d, int stage, const double *u0, const double *uin,
                      double *uout, const double *dt, int64_t dt_stride, const uint8_t *active,
                      double *lf_work, double *maxabs, int ghost_rows, psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (stage < 0 || stage > 3 || uin == nullptr || uout == nullptr) return PSK_E_INVALID;
  if (stage != 0 && dt == nullptr) return PSK_E_INVALID;
  if (stage >= 2 && u0 == nullptr) return PSK_E_INVALID;
  if (uout == uin) return PSK_E_INVALID;  // tiles read their neighbours' cells
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = lf_speed_pass(d, uin, 0, lf_work, st);
  if (rc != PSK_OK) return rc;
  StageParams p = make_params(d);
  p.uin = uin;
  p.u0 = u0;
  p.uout = uout;
  p.dt = dt;
  p.dt_stride = dt_stride;
  p.active = active;
  p.lf_speed = lf_work;
  p.maxabs = reinterpret_cast<unsigned long long *>(maxabs);
  p.stage = stage;
  return d->math == PSK_MATH_STRICT ? dispatch_scheme<true>(d, p, ghost_rows, st)
                                    : dispatch_scheme<false>(d, p, ghost_rows, st);
}

/* uout = ca u0 + cb uin + cc dt L(uin): the fused RHS with a general combine (timestepping.py:289-405) */
int psk_rhs_axpby(const psk_desc *d, const double *u0, const double *uin, double *uout, const double *dt,
                  int64_t dt_stride, double ca, double cb, double cc, double *lf_work, int ghost_rows,
                  psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (u0 == nullptr || uin == nullptr || uout == nullptr || dt == nullptr || uout == uin) return PSK_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = lf_speed_pass(d, uin, 0, lf_work, st);
  if (rc != PSK_OK) return rc;
  StageParams p = make_params(d);
  p.uin = uin;
  p.u0 = u0;
  p.uout = uout;
  p.dt = dt;
  p.dt_stride = dt_stride;
  p.lf_speed = lf_work;
  p.stage = 4;
  p.ca = ca; p.cb = cb; p.cc = cc;
  return d->math == PSK_MATH_STRICT ? dispatch_scheme<true>(d, p, ghost_rows, st)
                                    : dispatch_scheme<false>(d, p, ghost_rows, st);
}

/* psk_ssprk33_stage for the global Lax-Friedrichs flux on PERIODIC rows without the reduction pass: the speed
 * max |w| over all cells after the boundary condition (scalar.py:277) equals max |uin| over the interior there (the
 * ghost cells are copies), which the stage that PRODUCED uin has already reduced into its maxabs output. */
int psk_ssprk33_stage_lf(const psk_desc *d, int stage, const double *u0, const double *uin, double *uout,
                         const double *dt, int64_t dt_stride, const double *speed, double *maxabs,
                         psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (stage < 1 || stage > 3 || uin == nullptr || uout == nullptr || dt == nullptr || speed == nullptr) return PSK_E_INVALID;
  if (stage >= 2 && u0 == nullptr) return PSK_E_INVALID;
  if (uout == uin) return PSK_E_INVALID;
  if (d->equation != PSK_EQ_BURGERS || d->flux != PSK_FLUX_LAX_FRIEDRICHS || d->bc != PSK_BC_PERIODIC)
    return PSK_E_UNSUPPORTED;
  StageParams p = make_params(d);
  p.uin = uin;
  p.u0 = u0;
  p.uout = uout;
  p.dt = dt;
  p.dt_stride = dt_stride;
  p.lf_speed = speed;
  p.maxabs = reinterpret_cast<unsigned long long *>(maxabs);
  p.stage = stage;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d->math == PSK_MATH_STRICT ? dispatch_scheme<true>(d, p, 0, st) : dispatch_scheme<false>(d, p, 0, st);
}

int psk_ssprk33_step(const psk_desc *d, const double *u, double *uout, const double *dt, int64_t dt_stride,
                     const uint8_t *active, double *maxabs, psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (u == nullptr || uout == nullptr || dt == nullptr || uout == u) return PSK_E_INVALID;
  const bool aligned = (reinterpret_cast<uintptr_t>(u + d->g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(uout + d->g) % 16 == 0) && (d->ld % 2 == 0);
  // (global Lax-Friedrichs: the row-wide maximum of every stage lives in one thread-block cluster -- whole rows only)
  const bool flux_ok = d->flux == PSK_FLUX_RUSANOV || d->flux == PSK_FLUX_UPWIND || d->flux == PSK_FLUX_ENGQUIST_OSHER ||
                       (d->flux == PSK_FLUX_LAX_FRIEDRICHS && d->bc == PSK_BC_PERIODIC);
  // advection / continuity: periodic rows only (on slabs the velocity would need ghost cells of its own)
  const bool eq_ok = d->equation == PSK_EQ_BURGERS ? (flux_ok && d->nu == nullptr)
                                                   : (d->flux == PSK_FLUX_UPWIND && d->bc == PSK_BC_PERIODIC);
  if (!eq_ok || d->rec != PSK_REC_WENOJS53 || d->math != PSK_MATH_FAST || !aligned || g_step_variant == 0 ||
      !((d->bc == PSK_BC_PERIODIC && d->g >= 3) || (d->bc == PSK_BC_NONE && d->g >= 9)))
    return PSK_E_UNSUPPORTED;
  return launch_step_fused(d, u, uout, dt, dt_stride, active, maxabs, static_cast<cudaStream_t>(stream));
}

int psk_ssprk33_step_bc(const psk_desc *d, const double *u, double *uout, const double *dt, int64_t dt_stride,
                        const double *ghost3, const uint8_t *active, double *maxabs, double *k1_out, double *k2_out,
                        psk_stream_t stream) {
  if (d == nullptr) return PSK_E_INVALID;
  psk_desc d2 = *d;
  if (ghost3 != nullptr) d2.ghost = ghost3;  // (check_desc wants boundary data for Dirichlet rows)
  int rc = check_desc(&d2);
  if (rc != PSK_OK) return rc;
  if (u == nullptr || dt == nullptr || uout == u || ghost3 == nullptr) return PSK_E_INVALID;
  if ((k1_out == nullptr) != (k2_out == nullptr) || (uout == nullptr && k1_out == nullptr)) return PSK_E_INVALID;
  if (k1_out != nullptr && (active != nullptr || maxabs != nullptr || k1_out == u || k2_out == u || k1_out == k2_out))
    return PSK_E_INVALID;
  auto al = [&](const double *a) { return a == nullptr || reinterpret_cast<uintptr_t>(a + d->g) % 16 == 0; };
  const bool aligned = al(u) && al(uout) && al(k1_out) && al(k2_out) && (d->ld % 2 == 0);
  // (nu: Rusanov / Lax-Friedrichs with alpha != 1 take the viscosity of every face; no other flux has one)
  const bool burgers_ok = d->equation == PSK_EQ_BURGERS &&
                          (d->flux == PSK_FLUX_RUSANOV ||
                           (d->nu == nullptr && (d->flux == PSK_FLUX_UPWIND || d->flux == PSK_FLUX_ENGQUIST_OSHER)) ||
                           (d->flux == PSK_FLUX_LAX_FRIEDRICHS && d->bc == PSK_BC_DIRICHLET));
  const bool linear_ok = d->equation != PSK_EQ_BURGERS && d->flux == PSK_FLUX_UPWIND;
  if (!(burgers_ok || linear_ok) || d->rec != PSK_REC_WENOJS53 || d->math != PSK_MATH_FAST || !aligned ||
      g_step_variant == 0 || (d->bc != PSK_BC_DIRICHLET && d->bc != PSK_BC_NEUMANN) || d->g < 3 || d->n < d->g)
    return PSK_E_UNSUPPORTED;
  return launch_step_fused(&d2, u, uout, dt, dt_stride, active, maxabs, static_cast<cudaStream_t>(stream), k1_out,
                           k2_out, ghost3);
}

/* nsteps whole-step launches back to back, every state written straight onto the tape (no copies, no host
 * round trip): tape[0] = initial state (in), tape[m] = state after m steps (out). */
int psk_ssprk33_steps_tape(const psk_desc *d, double *tape, int64_t tape_stride, int nsteps, const double *dt_table,
                           const double *ghost_table, psk_stream_t stream) {
  if (d == nullptr || tape == nullptr || dt_table == nullptr || nsteps <= 0 || tape_stride <= 0) return PSK_E_INVALID;
  if ((d->bc == PSK_BC_DIRICHLET || d->bc == PSK_BC_NEUMANN) && ghost_table == nullptr) return PSK_E_INVALID;
  const int64_t ghost_block = d->ghost_ld != 0 ? static_cast<int64_t>(d->batch) * d->ghost_ld : 2 * d->g;
  for (int m = 0; m < nsteps; ++m) {
    const double *u = tape + static_cast<int64_t>(m) * tape_stride;
    double *un = tape + static_cast<int64_t>(m + 1) * tape_stride;
    int rc;
    if (d->bc == PSK_BC_DIRICHLET || d->bc == PSK_BC_NEUMANN)
      rc = psk_ssprk33_step_bc(d, u, un, dt_table + m, 0, ghost_table + static_cast<int64_t>(3) * m * ghost_block, nullptr,
                               nullptr, nullptr, nullptr, stream);
    else
      rc = psk_ssprk33_step(d, u, un, dt_table + m, 0, nullptr, nullptr, stream);
    if (rc != PSK_OK) return rc;  // (PSK_E_UNSUPPORTED can only come from the first step: nothing has run then)
  }
  return PSK_OK;
}

int psk_ssprk33_step_stages(const psk_desc *d, const double *u, double *k1, double *k2, double *uout,
                            const double *dt, int64_t dt_stride, psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (u == nullptr || k1 == nullptr || k2 == nullptr || dt == nullptr) return PSK_E_INVALID;
  if (uout == u || k1 == u || k2 == u || k1 == k2 || (uout != nullptr && (uout == k1 || uout == k2))) return PSK_E_INVALID;
  const bool aligned = (reinterpret_cast<uintptr_t>(u + d->g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(k1 + d->g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(k2 + d->g) % 16 == 0) &&
                       (uout == nullptr || reinterpret_cast<uintptr_t>(uout + d->g) % 16 == 0) && (d->ld % 2 == 0);
  if (d->equation != PSK_EQ_BURGERS || d->flux != PSK_FLUX_RUSANOV || d->rec != PSK_REC_WENOJS53 ||
      d->math != PSK_MATH_FAST || d->nu != nullptr || !aligned || g_step_variant == 0 ||
      !(d->bc == PSK_BC_PERIODIC && d->g >= 3))
    return PSK_E_UNSUPPORTED;
  return launch_step_fused(d, u, uout, dt, dt_stride, nullptr, nullptr, static_cast<cudaStream_t>(stream), k1, k2);
}

int psk_ssprk33_stage_p2p(const psk_desc *d, int stage, const double *u0, const double *uin,
                          double *uout, const double *dt, double *maxabs, const psk_halo_link *link,
                          psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (link == nullptr || stage < 1 || stage > 3 || uin == nullptr || uout == nullptr || dt == nullptr)
    return PSK_E_INVALID;
  if (stage >= 2 && u0 == nullptr) return PSK_E_INVALID;
  if (uout == uin || link->timeout_ns <= 0) return PSK_E_INVALID;
  if ((link->peer_lo != nullptr && link->flag_lo == nullptr) || (link->peer_hi != nullptr && link->flag_hi == nullptr))
    return PSK_E_INVALID;
  // the fused form exists for the hot configuration only; callers fall back to
  // psk_halo_wait -> psk_ssprk33_stage -> psk_halo_push otherwise
  const bool aligned = (reinterpret_cast<uintptr_t>(uin + d->g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(uout + d->g) % 16 == 0) &&
                       (u0 == nullptr || reinterpret_cast<uintptr_t>(u0 + d->g) % 16 == 0);
  if (d->equation != PSK_EQ_BURGERS || d->flux != PSK_FLUX_RUSANOV || d->rec != PSK_REC_WENOJS53 ||
      d->math != PSK_MATH_FAST || d->bc != PSK_BC_NONE || d->batch != 1 || d->g != 3 || d->nu != nullptr ||
      d->n % 4 != 0 || d->n < 8 || !aligned)
    return PSK_E_UNSUPPORTED;
  FastParams q{};
  q.uin = uin; q.u0 = u0; q.uout = uout; q.dt = dt;
  q.maxabs = reinterpret_cast<unsigned long long *>(maxabs);
  q.bc = make_bc_view(d);
  q.ld = d->ld;
  q.coef = (1.0 / d->dx) / FluxScale<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV>::value;
  q.eps9 = d->eps * (1.0 / 9.0);
  q.dt_stride = 0;
  q.chunks_per_row = (d->n + 119) / 120;
  HaloLink h{};
  h.wait_lo = reinterpret_cast<const long long *>(link->wait_lo);
  h.wait_hi = reinterpret_cast<const long long *>(link->wait_hi);
  h.wait_epoch = link->wait_epoch;
  h.peer_lo = link->peer_lo;
  h.peer_hi = link->peer_hi;
  h.flag_lo = reinterpret_cast<long long *>(link->flag_lo);
  h.flag_hi = reinterpret_cast<long long *>(link->flag_hi);
  h.timeout_ns = static_cast<unsigned long long>(link->timeout_ns);
  h.timed_out = link->timed_out;
  if ((link->epoch_in == nullptr) != (link->epoch_out == nullptr) ||
      (link->epoch_in != nullptr && link->epoch_in == link->epoch_out))
    return PSK_E_INVALID;
  h.epoch_in = reinterpret_cast<const long long *>(link->epoch_in);
  h.epoch_out = reinterpret_cast<long long *>(link->epoch_out);
  // 5 warps per CTA: 6 CTAs (30 warps) per SM at 64 registers, the shape the plain launcher
  // picks for long single rows
  const int wpc = q.chunks_per_row < 5 ? q.chunks_per_row : 5;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define PSK_P2P_LAUNCH(STAGE)                                                        \
  if (maxabs != nullptr)                                                             \
    stage_warp_fast_p2p_kernel<STAGE, true><<<gx, wpc * 32, 0, st>>>(q, h);          \
  else                                                                               \
    stage_warp_fast_p2p_kernel<STAGE, false><<<gx, wpc * 32, 0, st>>>(q, h)
  switch (stage) {
    case 1: PSK_P2P_LAUNCH(1); break;
    case 2: PSK_P2P_LAUNCH(2); break;
    default: PSK_P2P_LAUNCH(3);
  }
#undef PSK_P2P_LAUNCH
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

int psk_ssprk33_step_p2p(const psk_desc *d, const double *u, double *uout, const double *dt, double *maxabs,
                         const psk_halo_link *link, psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (link == nullptr || u == nullptr || uout == nullptr || dt == nullptr || uout == u || link->timeout_ns <= 0)
    return PSK_E_INVALID;
  if ((link->peer_lo != nullptr && link->flag_lo == nullptr) || (link->peer_hi != nullptr && link->flag_hi == nullptr))
    return PSK_E_INVALID;
  if (link->epoch_in != nullptr || link->epoch_out != nullptr) return PSK_E_UNSUPPORTED;  // no graph replay here
  const bool aligned = (reinterpret_cast<uintptr_t>(u + d->g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(uout + d->g) % 16 == 0);
  const bool flux_ok = d->flux == PSK_FLUX_RUSANOV || d->flux == PSK_FLUX_UPWIND || d->flux == PSK_FLUX_ENGQUIST_OSHER;
  constexpr int kEmit = StepGeometry<6>::kEmit;
  // each end chunk must hold the slab's 9 edge cells AND be the only one whose window reaches the ghost cells on
  // its side, i.e. hold at least 10 cells: if the plain chunk grid leaves fewer in the last chunk, the grid is
  // shifted left by half a chunk (first chunk 86 cells, last chunk r + 86)
  const int rem = d->n % kEmit;
  const int shift = (rem == 0 || rem >= 10) ? 0 : kEmit / 2;
  const int chunks = (d->n + shift + kEmit - 1) / kEmit;
  const bool one_edge_chunk = d->n >= 2 * kEmit;
  if (d->equation != PSK_EQ_BURGERS || !flux_ok || d->rec != PSK_REC_WENOJS53 || d->math != PSK_MATH_FAST ||
      d->nu != nullptr || d->bc != PSK_BC_NONE || d->batch != 1 || d->g != 9 || d->n < 18 || !aligned || !one_edge_chunk)
    return PSK_E_UNSUPPORTED;
  StepParams q{};
  q.u = u; q.uout = uout; q.dt = dt;
  q.maxabs = reinterpret_cast<unsigned long long *>(maxabs);
  q.ld = d->ld;
  q.coef = (1.0 / d->dx) / (d->flux == PSK_FLUX_RUSANOV ? 4.0 : 2.0);
  q.eps9 = d->eps * (1.0 / 9.0);
  q.n = d->n; q.g = d->g; q.bc_none = 1;
  q.chunks_per_row = chunks;
  q.shift = shift;
  HaloLink h{};
  h.wait_lo = reinterpret_cast<const long long *>(link->wait_lo);
  h.wait_hi = reinterpret_cast<const long long *>(link->wait_hi);
  h.wait_epoch = link->wait_epoch;
  h.peer_lo = link->peer_lo;
  h.peer_hi = link->peer_hi;
  h.flag_lo = reinterpret_cast<long long *>(link->flag_lo);
  h.flag_hi = reinterpret_cast<long long *>(link->flag_hi);
  h.timeout_ns = static_cast<unsigned long long>(link->timeout_ns);
  h.timed_out = link->timed_out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->flux == PSK_FLUX_UPWIND) return launch_step_p2p<PSK_FLUX_UPWIND>(q, h, maxabs != nullptr, st);
  if (d->flux == PSK_FLUX_ENGQUIST_OSHER) return launch_step_p2p<PSK_FLUX_ENGQUIST_OSHER>(q, h, maxabs != nullptr, st);
  return launch_step_p2p<PSK_FLUX_RUSANOV>(q, h, maxabs != nullptr, st);
}

int psk_apply_operator(const psk_desc *d, const double *u, double *rhs, double *lf_work,
                       psk_stream_t stream) {
  return psk_ssprk33_stage(d, 0, nullptr, u, rhs, nullptr, 0, nullptr, lf_work, nullptr, 1,
                           stream);
}

int psk_step_control(int32_t batch, double theta, double cfl_scale, double tfinal,
                     const double *maxabs, const double *t, double *t_next, double *dt,
                     uint8_t *active, int32_t *nonfinite, psk_stream_t stream) {
  if (batch <= 0 || maxabs == nullptr || t == nullptr || t_next == nullptr || dt == nullptr ||
      active == nullptr || nonfinite == nullptr)
    return PSK_E_INVALID;
  step_control_kernel<<<(batch + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      batch, theta, cfl_scale, tfinal, maxabs, t, t_next, dt, active, nonfinite);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

}  // extern "C"

