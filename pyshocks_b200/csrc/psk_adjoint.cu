// psk_adjoint.cu -- hand-derived discrete adjoint of the fused stage (sm_100a, fp64).
//
// The reference builds the dense Jacobian of one SSPRK33 step with jax.jacfwd and applies
// its transpose (timestepping.py:174, :205-206): O(nx^2) work and memory per step.  Here the
// transposed stencil is applied matrix-free, O(nx), one launch per stage:
//
//   out = c_acc acc + c_acc2 acc2 + c_v v + c_g dt J_L(x)^T v
//
// with L = apply_operator (boundary condition included) and J_L evaluated on ALL nx rows,
// ghost rows included, exactly as the reference's full-array `advance` is differentiated
// (SURVEY.md 3.3 "ghost rows" quirk).
//
// Kernel structure (adjoint_tile_kernel): the transpose of the forward tile kernel.
//   * a CTA stages w = BC(x) and v for its cells plus halo in shared memory;
//   * each thread owns R consecutive cells; pass 1 recomputes their face values, neighbours
//     exchange one value each way, and the R + 1 face-flux derivatives give the cotangents
//     of every face value (g_ur, g_ul) and the direct cell terms;
//   * pass 2 pushes (g_ur, g_ul) through the WENO weights (psk_math.cuh, weno53_pair_vjp)
//     and accumulates the 5-point scatter in registers; only the two-cell spill on each side
//     goes through shared memory;
//   * the first and last thread of a CTA are halo threads (they compute, they do not store),
//     so no atomics are needed between tiles;
//   * cotangents that land on ghost cells are parked in a per-row spill area and folded back
//     onto the cells the boundary condition copied them from by a tiny second kernel (the
//     transpose of apply_boundary), which also distributes the global Lax-Friedrichs speed
//     cotangent onto the arg-max cells.
#include <cuda_runtime.h>

#include <cstdint>

#include "psk_common.cuh"
#include "psk_math.cuh"
#include "psk_adjoint_math.cuh"
#include "psk_adjoint_kernels.cuh"

namespace psk {


constexpr int kAdjHalo = 3;

template <int R>
__host__ __device__ __forceinline__ int adj_pad(int e) {
  return e + e / R;
}

template <int EQ, int FLUX, int REC, int R>
__global__ void __launch_bounds__(256)
adjoint_tile_kernel(const AdjParams p) {
  extern __shared__ double smem[];
  const int nthreads = blockDim.x;
  const int out_cells = (nthreads - 2) * R;
  const int row = blockIdx.x / p.tiles_per_row;
  const int tile = blockIdx.x - row * p.tiles_per_row;
  const int nx = p.bc.nx, g = p.bc.g;
  const int S = tile * out_cells - R;  // array index of the first cell of (halo) thread 0
  const double *__restrict__ xrow = p.x + static_cast<int64_t>(row) * p.ld;
  const double *__restrict__ vrow = p.v + static_cast<int64_t>(row) * p.ld;

  // shared: w tile | v tile | XL | XR | SL (2 per thread) | SR (2 per thread) | warp sums
  const int elems = nthreads * R + 2 * kAdjHalo;
  double *tw = smem;
  double *tv = tw + adj_pad<R>(elems) + 1;
  double *xl = tv + adj_pad<R>(elems) + 1;
  double *xr = xl + nthreads + 1;
  double *sl = xr + nthreads + 1;
  double *sr = sl + 2 * nthreads;
  double *wsum = sr + 2 * nthreads;

  for (int e = threadIdx.x; e < elems; e += nthreads) {
    const int i = S - kAdjHalo + e;
    tw[adj_pad<R>(e)] = load_w(p.bc, xrow, row, i);
    tv[adj_pad<R>(e)] = (i >= 0 && i < nx) ? vrow[i] : 0.0;
  }
  __syncthreads();

  const int t = threadIdx.x;
  const int c0 = S + R * t;
  double w[R + 2 * kAdjHalo], vw[R + 2];
  {
    const double *bw = tw + (R + 1) * t;
    const double *bv = tv + (R + 1) * t;
#pragma unroll
    for (int k = 0; k < R + 2 * kAdjHalo; ++k) w[k] = bw[k + k / R];
#pragma unroll
    for (int k = 0; k < R + 2; ++k) vw[k] = bv[(k + 2) + (k + 2) / R];  // v[c0 - 1 + k]
  }

  // ---- pass 1: face values of the owned cells, neighbour exchange
  // (ESWENO32: tau is zero in the first and the last cell of the array, weno.py:292)
  auto tz = [&](int i) { return REC == PSK_REC_ESWENO32 && (i <= 0 || i >= nx - 1); };
  double ul[R], ur[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = r + kAdjHalo;
    Weno5Pair o = reconstruct_cell<REC, false>(w[m - 2], w[m - 1], w[m], w[m + 1], w[m + 2], p.eps, tz(c0 + r));
    ul[r] = o.ul;
    ur[r] = o.ur;
  }
  xl[t] = ul[0];
  xr[t + 1] = ur[R - 1];
  if (t == 0) xr[0] = reconstruct_cell<REC, false>(w[0], w[1], w[2], w[3], w[4], p.eps, tz(c0 - 1)).ur;
  if (t == nthreads - 1)
    xl[nthreads] =
        reconstruct_cell<REC, false>(w[R + 1], w[R + 2], w[R + 3], w[R + 4], w[R + 5], p.eps, tz(c0 + R)).ul;
  __syncthreads();
  // the Burgers ESWENO32 scheme: omega_0 of the cells c0 - 1 .. c0 + R for the dissipative flux of their faces
  double om[R + 2], gom[R];
#pragma unroll
  for (int r = 0; r < R; ++r) gom[r] = 0.0;
  if (FLUX == PSK_FLUX_ESWENO) {
#pragma unroll
    for (int k = 0; k < R + 2; ++k) om[k] = esweno32_cell<false>(w[k + 1], w[k + 2], w[k + 3], p.eps, tz(c0 - 1 + k)).om0;
  }
  const double ur_left = xr[t];
  const double ul_right = xl[t + 1];

  // ---- face-flux derivatives: cotangents of the face values and direct cell terms
  const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? p.speed[row] : 0.0;
  const bool writer = (t >= 1 && t <= nthreads - 2);
  double gur[R + 1], gul[R + 1];  // gur[f] belongs to cell c0 + f - 1, gul[f] to cell c0 + f
  double dir[R];
#pragma unroll
  for (int r = 0; r < R; ++r) dir[r] = 0.0;
  double ga_part = 0.0;
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const int k = c0 + f;  // face between cells k - 1 and k
    const bool valid = (k >= 1 && k <= nx - 1);
    const double urj = (f == 0) ? ur_left : ur[f - 1];
    const double ulp = (f == R) ? ul_right : ul[f];
    double nu = 1.0, arj = 0.0, alp = 0.0, ck = 1.0, ckm1 = 1.0;
    if (valid) {
      if ((FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) && p.nu != nullptr)
        nu = p.nu[k - 1];
      if (EQ != PSK_EQ_BURGERS) {
        arj = p.vel_r[k - 1];
        alp = p.vel_l[k];
      }
      if (EQ == PSK_EQ_ADVECTION) {
        ck = p.vel[k];
        ckm1 = p.vel[k - 1];
      }
    }
    // L[i] = -c_i (F[i+1] - F[i]) / dx  ->  gF[k] = (c_k v[k] - c_{k-1} v[k-1]) / dx
    const double gF = valid ? (ck * vw[f + 1] - ckm1 * vw[f]) * p.invdx : 0.0;
    const FaceGrad fg = face_flux_grad<EQ, FLUX>(urj, ulp, w[f + kAdjHalo - 1], w[f + kAdjHalo],
                                                 speed, nu, arj, alp);
    gur[f] = gF * fg.d_ur;
    gul[f] = gF * fg.d_ul;
    if (f >= 1) dir[f - 1] = fma(gF, fg.d_wj, dir[f - 1]);
    if (f <= R - 1) dir[f] = fma(gF, fg.d_wp, dir[f]);
    if (FLUX == PSK_FLUX_LAX_FRIEDRICHS && f >= 1 && writer) ga_part = fma(gF, fg.d_speed, ga_part);
    if (FLUX == PSK_FLUX_ESWENO) {  // + g(om_j, om_p, w_j, w_p) (burgers/schemes.py:243-256)
      const EsGnumGrad eg = esweno_gnum_grad(om[f], om[f + 1], w[f + kAdjHalo - 1], w[f + kAdjHalo], p.delta);
      if (f >= 1) {
        dir[f - 1] = fma(gF, eg.d_wj, dir[f - 1]);
        gom[f - 1] = fma(gF, eg.d_omj, gom[f - 1]);
      }
      if (f <= R - 1) {
        dir[f] = fma(gF, eg.d_wp, dir[f]);
        gom[f] = fma(gF, eg.d_omp, gom[f]);
      }
    }
  }

  // ---- pass 2: through the reconstruction, 5-point scatter kept in registers
  double accw[R + 4];  // cotangents of cells c0 - 2 .. c0 + R + 1
#pragma unroll
  for (int k = 0; k < R + 4; ++k) accw[k] = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = r + kAdjHalo;
    const Weno5Vjp d = reconstruct_cell_vjp<REC>(w[m - 2], w[m - 1], w[m], w[m + 1], w[m + 2], p.eps,
                                                 gur[r + 1], gul[r], tz(c0 + r), gom[r]);
#pragma unroll
    for (int q = 0; q < 5; ++q) accw[r + q] += d.d[q];
  }
  sl[2 * t] = accw[0];
  sl[2 * t + 1] = accw[1];
  sr[2 * t] = accw[R + 2];
  sr[2 * t + 1] = accw[R + 3];
  __syncthreads();

  if (writer) {
    const double cgdt =
        p.c_g * (p.dt != nullptr ? p.dt[static_cast<int64_t>(row) * p.dt_stride] : 1.0);
    const int64_t base = static_cast<int64_t>(row) * p.ld;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = c0 + r;
      if (i < 0 || i >= nx) continue;
      double gi = accw[r + 2] + dir[r];
      if (r < 2) gi += sr[2 * (t - 1) + r];
      if (r >= R - 2) gi += sl[2 * (t + 1) + (r - (R - 2))];
      double lin = p.c_v * vw[r + 1];
      if (p.acc != nullptr) lin = fma(p.c_acc, p.acc[base + i], lin);
      if (p.acc2 != nullptr) lin = fma(p.c_acc2, p.acc2[base + i], lin);
      const bool ghost = (p.bc.bc != PSK_BC_NONE) && (i < g || i >= nx - g);
      if (ghost) {
        p.out[base + i] = lin;  // ghost cells of x do not influence L: (J^T v)[ghost] = 0
        p.gspill[static_cast<int64_t>(row) * 2 * g + (i < g ? i : i - p.bc.n)] = gi;
      } else {
        p.out[base + i] = fma(cgdt, gi, lin);
      }
    }
  }

  if (FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ga_part += __shfl_xor_sync(0xffffffffu, ga_part, off);
    const int lane = t & 31, wid = t >> 5;
    if (lane == 0) wsum[wid] = ga_part;
    __syncthreads();
    if (t == 0) {
      double s = 0.0;
      for (int k = 0; k < (nthreads + 31) / 32; ++k) s += wsum[k];
      atomicAdd(p.ga + row, s);
    }
  }
}

// ---------------------------------------------------------------------------
// Warp form of the adjoint stage for WENO-JS5 (the transposed counterpart of
// stage_warp_fast_kernel): one warp covers 128 consecutive cells, lanes 0 and 31 are halo lanes,
// all neighbour traffic (3-cell halos of w, one cotangent value, one face value and the
// two-cell scatter spill each way) goes through shuffles; no shared memory, no barriers.
// Because a cell's cotangent reaches two cells further than its value did in the forward
// sweep, the two edge lanes fetch one extra cell each.
template <int EQ, int FLUX>
__global__ void __launch_bounds__(256, 2)
adjoint_warp_kernel(const AdjParams p, int chunks_per_row) {
  constexpr int R = 4;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kOut = 30 * R;
  const int lane = threadIdx.x & 31;
  // chunk -1 covers the left ghost cells (interior coordinates -g .. -1)
  const int chunk = static_cast<int>(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) - 1;
  if (chunk >= chunks_per_row) return;
  const int row = blockIdx.y + blockIdx.z * gridDim.y;
  const int g = p.bc.g, n = p.bc.n, nx = p.bc.nx;
  const int c0 = chunk * kOut - R + R * lane;  // interior coordinates; array index = g + c0
  const int64_t base = static_cast<int64_t>(row) * p.ld;
  const int64_t off = base + g + c0;
  const bool inside = (c0 >= 0) && (c0 + R <= n);
  const bool emit = (lane >= 1) && (lane <= 30);
  const double *__restrict__ xrow = p.x + base;
  const double *__restrict__ vrow = p.v + base;

  // ---- forward state window and cotangent window
  double w[R + 6], vc[R + 2];
  if (inside) {
    const double2 q0 = *reinterpret_cast<const double2 *>(p.x + off);
    const double2 q1 = *reinterpret_cast<const double2 *>(p.x + off + 2);
    w[3] = q0.x; w[4] = q0.y; w[5] = q1.x; w[6] = q1.y;
    const double2 r0 = *reinterpret_cast<const double2 *>(p.v + off);
    const double2 r1 = *reinterpret_cast<const double2 *>(p.v + off + 2);
    vc[1] = r0.x; vc[2] = r0.y; vc[3] = r1.x; vc[4] = r1.y;
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = g + c0 + r;
      w[3 + r] = load_w(p.bc, xrow, row, i);
      vc[1 + r] = (i >= 0 && i < nx) ? vrow[i] : 0.0;
    }
  }
  double extra = 0.0;  // lane 0: cell c0 - 1, lane 31: cell c0 + R
  if (lane == 0) extra = load_w(p.bc, xrow, row, g + c0 - 1);
  if (lane == 31) extra = load_w(p.bc, xrow, row, g + c0 + R);
  w[0] = __shfl_up_sync(kFull, w[4], 1);
  w[1] = __shfl_up_sync(kFull, w[5], 1);
  w[2] = __shfl_up_sync(kFull, w[6], 1);
  w[7] = __shfl_down_sync(kFull, w[3], 1);
  w[8] = __shfl_down_sync(kFull, w[4], 1);
  w[9] = __shfl_down_sync(kFull, w[5], 1);
  if (lane == 0) w[2] = extra;
  if (lane == 31) w[7] = extra;
  vc[0] = __shfl_up_sync(kFull, vc[R], 1);
  vc[R + 1] = __shfl_down_sync(kFull, vc[1], 1);

  // ---- pass 1: face values
  const double eps9 = p.eps * (1.0 / 9.0);
  double t[R + 5], pq[R + 4];
#pragma unroll
  for (int k = 0; k < R + 5; ++k) t[k] = (1.0 / 6.0) * (w[k + 1] - w[k]);
#pragma unroll
  for (int k = 0; k < R + 4; ++k) {
    const double dd = t[k + 1] - t[k];
    pq[k] = fma((13.0 / 3.0) * dd, dd, eps9);
  }
  double ul[R], ur[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = r + 3;
    const Weno5Pair o = weno53_pair_lean(w[m], t[m - 2], t[m - 1], t[m], t[m + 1], pq[m - 2], pq[m - 1], pq[m]);
    ul[r] = o.ul;
    ur[r] = o.ur;
  }
  const double ur_left = __shfl_up_sync(kFull, ur[R - 1], 1);
  const double ul_right = __shfl_down_sync(kFull, ul[0], 1);

  // ---- face-flux derivatives
  const double speed = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? p.speed[row] : 0.0;
  const bool has_nu = (FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) && (p.nu != nullptr);
  double gur[R + 1], gul[R + 1], dir[R];
#pragma unroll
  for (int r = 0; r < R; ++r) dir[r] = 0.0;
  double ga_part = 0.0;
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const int k = g + c0 + f;  // array index of the face: between cells k - 1 and k
    const bool valid = (k >= 1 && k <= nx - 1);
    const double urj = (f == 0) ? ur_left : ur[f - 1];
    const double ulp = (f == R) ? ul_right : ul[f];
    double nu = 1.0, arj = 0.0, alp = 0.0, ck = 1.0, ckm1 = 1.0;
    if (valid) {
      if (has_nu) nu = p.nu[k - 1];
      if (EQ != PSK_EQ_BURGERS) {
        arj = p.vel_r[k - 1];
        alp = p.vel_l[k];
      }
      if (EQ == PSK_EQ_ADVECTION) {
        ck = p.vel[k];
        ckm1 = p.vel[k - 1];
      }
    }
    const double gF = valid ? (ck * vc[f + 1] - ckm1 * vc[f]) * p.invdx : 0.0;
    const FaceGrad fg = face_flux_grad<EQ, FLUX>(urj, ulp, w[f + 2], w[f + 3], speed, nu, arj, alp);
    gur[f] = gF * fg.d_ur;
    gul[f] = gF * fg.d_ul;
    if (f >= 1) dir[f - 1] = fma(gF, fg.d_wj, dir[f - 1]);
    if (f <= R - 1) dir[f] = fma(gF, fg.d_wp, dir[f]);
    if (FLUX == PSK_FLUX_LAX_FRIEDRICHS && f >= 1 && emit) ga_part = fma(gF, fg.d_speed, ga_part);
  }

  // ---- pass 2: through the reconstruction; 5-point scatter in registers
  double accw[R + 4];
#pragma unroll
  for (int k = 0; k < R + 4; ++k) accw[k] = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = r + 3;
    const Weno5Vjp d = weno53_pair_vjp_sixths(t[m - 2], t[m - 1], t[m], t[m + 1], pq[m - 2], pq[m - 1], pq[m],
                                              gur[r + 1], gul[r]);
#pragma unroll
    for (int q = 0; q < 5; ++q) accw[r + q] += d.d[q];
  }
  // spill: my contributions to the neighbour lanes' cells
  const double from_right0 = __shfl_down_sync(kFull, accw[0], 1);  // next lane -> my cell R-2
  const double from_right1 = __shfl_down_sync(kFull, accw[1], 1);  // next lane -> my cell R-1
  const double from_left0 = __shfl_up_sync(kFull, accw[R + 2], 1);  // previous lane -> my cell 0
  const double from_left1 = __shfl_up_sync(kFull, accw[R + 3], 1);  // previous lane -> my cell 1

  if (emit) {
    const double cgdt = p.c_g * (p.dt != nullptr ? p.dt[static_cast<int64_t>(row) * p.dt_stride] : 1.0);
    double gi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) gi[r] = accw[r + 2] + dir[r];
    gi[0] += from_left0;
    gi[1] += from_left1;
    gi[R - 2] += from_right0;
    gi[R - 1] += from_right1;
    double lin[R];
    if (inside) {
#pragma unroll
      for (int r = 0; r < R; ++r) lin[r] = p.c_v * vc[r + 1];
      if (p.acc != nullptr) {
        const double2 a0 = *reinterpret_cast<const double2 *>(p.acc + off);
        const double2 a1 = *reinterpret_cast<const double2 *>(p.acc + off + 2);
        lin[0] = fma(p.c_acc, a0.x, lin[0]); lin[1] = fma(p.c_acc, a0.y, lin[1]);
        lin[2] = fma(p.c_acc, a1.x, lin[2]); lin[3] = fma(p.c_acc, a1.y, lin[3]);
      }
      if (p.acc2 != nullptr) {
        const double2 a0 = *reinterpret_cast<const double2 *>(p.acc2 + off);
        const double2 a1 = *reinterpret_cast<const double2 *>(p.acc2 + off + 2);
        lin[0] = fma(p.c_acc2, a0.x, lin[0]); lin[1] = fma(p.c_acc2, a0.y, lin[1]);
        lin[2] = fma(p.c_acc2, a1.x, lin[2]); lin[3] = fma(p.c_acc2, a1.y, lin[3]);
      }
      // inside => interior cells only (no ghost among them)
      *reinterpret_cast<double2 *>(p.out + off) = make_double2(fma(cgdt, gi[0], lin[0]), fma(cgdt, gi[1], lin[1]));
      *reinterpret_cast<double2 *>(p.out + off + 2) = make_double2(fma(cgdt, gi[2], lin[2]), fma(cgdt, gi[3], lin[3]));
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = g + c0 + r;
        if (i < 0 || i >= nx) continue;
        double l = p.c_v * vc[r + 1];
        if (p.acc != nullptr) l = fma(p.c_acc, p.acc[base + i], l);
        if (p.acc2 != nullptr) l = fma(p.c_acc2, p.acc2[base + i], l);
        const bool ghost = (p.bc.bc != PSK_BC_NONE) && (i < g || i >= nx - g);
        if (ghost) {
          p.out[base + i] = l;
          p.gspill[static_cast<int64_t>(row) * 2 * g + (i < g ? i : i - n)] = gi[r];
        } else {
          p.out[base + i] = fma(cgdt, gi[r], l);
        }
      }
    }
  }
  if (FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ga_part += __shfl_xor_sync(kFull, ga_part, o);
    if (lane == 0) atomicAdd(p.ga + row, ga_part);
  }
}


template <int MINB, int VAR, int FLUX = PSK_FLUX_RUSANOV, bool NU = false>
int launch_adjoint_lean(const AdjParams &p, int batch, cudaStream_t st) {
  const int chunks = (p.bc.n + p.bc.g + 119) / 120;  // chunks 0 .. chunks-1 plus chunk -1
  const int total = chunks + 1;
  const int wpc = total < 4 ? total : 4;
  const unsigned gx = static_cast<unsigned>((total + wpc - 1) / wpc);
  unsigned gy, gz;
  if (!split_rows(batch, gy, gz)) return PSK_E_UNSUPPORTED;
  const dim3 grid(gx, gy, gz);
  adjoint_lean_kernel<MINB, VAR, FLUX, NU><<<grid, wpc * 32, 0, st>>>(p, chunks);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

template <int EQ, int FLUX>
int launch_adjoint_warp(const AdjParams &p, int batch, cudaStream_t st) {
  const int chunks = (p.bc.n + p.bc.g + 119) / 120;  // chunks 0 .. chunks-1 plus chunk -1
  const int total = chunks + 1;
  int wpc = 8, best_waste = 1 << 30;
  for (int w = 8; w >= 4; --w) {
    const int waste = ((total + w - 1) / w) * w - total;
    if (waste < best_waste) { best_waste = waste; wpc = w; }
  }
  if (total < wpc) wpc = total;
  const unsigned gx = static_cast<unsigned>((total + wpc - 1) / wpc);
  unsigned gy, gz;
  if (!split_rows(batch, gy, gz)) return PSK_E_UNSUPPORTED;
  const dim3 grid(gx, gy, gz);
  adjoint_warp_kernel<EQ, FLUX><<<grid, wpc * 32, 0, st>>>(p, chunks);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

// transpose of apply_boundary + distribution of the Lax-Friedrichs speed cotangent; one CTA per row
template <bool LF>
__global__ void adjoint_boundary_kernel(const AdjParams p) {
  const int row = blockIdx.x;
  const int nx = p.bc.nx, g = p.bc.g, n = p.bc.n;
  const double *__restrict__ xrow = p.x + static_cast<int64_t>(row) * p.ld;
  double *orow = p.out + static_cast<int64_t>(row) * p.ld;
  const double cgdt =
      p.prescaled ? 1.0 : p.c_g * (p.dt != nullptr ? p.dt[static_cast<int64_t>(row) * p.dt_stride] : 1.0);
  // where apply_boundary copied ghost cell i from (-1: from nowhere)
  auto source = [&](int i) -> int {
    if (i >= g && i < nx - g) return i;
    switch (p.bc.bc) {
      case PSK_BC_PERIODIC: return i < g ? i + n : i - n;
      case PSK_BC_NEUMANN: return i < g ? 2 * g - 1 - i : 2 * (nx - g) - 1 - i;
      case PSK_BC_DIRICHLET: return -1;
      default: return i;
    }
  };
  if (p.bc.bc == PSK_BC_PERIODIC || p.bc.bc == PSK_BC_NEUMANN) {
    for (int k = threadIdx.x; k < 2 * g; k += blockDim.x) {
      const int i = k < g ? k : nx - 2 * g + k;
      atomicAdd(orow + source(i), cgdt * p.gspill[static_cast<int64_t>(row) * 2 * g + k]);
    }
  }
  if (LF) {
    // jnp.max(jnp.abs(w)): the cotangent is shared equally between the tied arg-max cells
    const double speed = p.speed[row];
    // the lean kernel counted the arg-max cells and kept the index of one: a single one needs no scan of the row
    if (p.amax != nullptr && p.amax[2 * row] == 1u) {
      if (threadIdx.x == 0) {
        const int i = static_cast<int>(p.amax[2 * row + 1]);
        const int src = source(i);
        if (src >= 0) atomicAdd(orow + src, cgdt * p.ga[row] * sign0(load_w(p.bc, xrow, row, i)));
      }
      return;
    }
    __shared__ int count;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < nx; i += blockDim.x)
      mine += (fabs(load_w(p.bc, xrow, row, i)) == speed) ? 1 : 0;
    if (mine) atomicAdd(&count, mine);
    __syncthreads();
    const double share = cgdt * p.ga[row] / static_cast<double>(count > 0 ? count : 1);
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      const double wi = load_w(p.bc, xrow, row, i);
      if (fabs(wi) == speed) {
        const int src = source(i);
        if (src >= 0) atomicAdd(orow + src, share * sign0(wi));
      }
    }
  }
}

// 0: lean warp kernel for the hot configuration, warp kernel for the other WENO-JS5 cases;
// 1: tile kernel always; 2: warp kernel (no lean kernel); 3, 5: lean kernel compiled for 3 / 5 CTAs
// of 128 threads per SM (register budgets 168 / 96; the default is 4 CTAs, 128 registers, linear part
// parked in shared memory, state of cell 3 recomputed); 6: 4 CTAs without those two measures
static int g_adjoint_variant = 0;

template <int EQ, int FLUX, int REC>
int launch_adjoint(const AdjParams &p0, int batch, cudaStream_t st) {
  constexpr int R = 4;
  AdjParams p = p0;
  constexpr bool LFX = (FLUX == PSK_FLUX_LAX_FRIEDRICHS);
  const bool aligned = (reinterpret_cast<uintptr_t>(p.x + p.bc.g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p.v + p.bc.g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p.out + p.bc.g) % 16 == 0) &&
                       (p.acc == nullptr || reinterpret_cast<uintptr_t>(p.acc + p.bc.g) % 16 == 0) &&
                       (p.acc2 == nullptr || reinterpret_cast<uintptr_t>(p.acc2 + p.bc.g) % 16 == 0) &&
                       (p.ld % 2 == 0);
  if (REC == PSK_REC_WENOJS53 && g_adjoint_variant != 1 && aligned && p.bc.g == 3 &&
      (batch <= 65535 || batch % 65535 == 0 || batch % 32768 == 0)) {
    int rc;
    if (EQ == PSK_EQ_BURGERS && FLUX == PSK_FLUX_RUSANOV && p.nu == nullptr && g_adjoint_variant != 2) {
      p.prescaled = 1;
      switch (g_adjoint_variant) {
        case 3: rc = launch_adjoint_lean<3, 0>(p, batch, st); break;
        case 5: rc = launch_adjoint_lean<5, 3>(p, batch, st); break;
        case 6: rc = launch_adjoint_lean<4, 0>(p, batch, st); break;
        default: rc = launch_adjoint_lean<4, 3>(p, batch, st); break;
      }
    } else if (EQ == PSK_EQ_BURGERS && (FLUX == PSK_FLUX_RUSANOV || LFX) && g_adjoint_variant != 2) {
      // the lean kernel with the viscosity of every face (alpha != 1) and / or the row's global speed (Lax-Friedrichs:
      // the scheme of the reference's burgers-adjoint driver); 0.41 ms against the general warp kernel's 0.79 ms per
      // stage of config 5
      p.prescaled = 1;
      constexpr int kFl = LFX ? PSK_FLUX_LAX_FRIEDRICHS : PSK_FLUX_RUSANOV;
      rc = p.nu != nullptr ? launch_adjoint_lean<4, 3, kFl, true>(p, batch, st)
                           : launch_adjoint_lean<4, 3, kFl, false>(p, batch, st);
    } else {
      rc = launch_adjoint_warp<EQ, FLUX>(p, batch, st);
    }
    if (rc != PSK_OK) return rc;
    if (LFX || p.bc.bc == PSK_BC_PERIODIC || p.bc.bc == PSK_BC_NEUMANN) {
      adjoint_boundary_kernel<LFX><<<batch, LFX ? 128 : 32, 0, st>>>(p);
      PSK_CUDA_OK(cudaGetLastError());
    }
    return PSK_OK;
  }
  const int nx = p.bc.nx;
  int threads = (nx + R - 1) / R + 2;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  const int out_cells = (threads - 2) * R;
  p.tiles_per_row = (nx + out_cells - 1) / out_cells;
  const int elems = threads * R + 2 * kAdjHalo;
  const size_t smem =
      sizeof(double) * (2 * (adj_pad<R>(elems) + 1) + 2 * (threads + 1) + 4 * threads + 8);
  const long long blocks = static_cast<long long>(p.tiles_per_row) * batch;
  if (blocks > 2147483647LL) return PSK_E_INVALID;
  adjoint_tile_kernel<EQ, FLUX, REC, R><<<static_cast<unsigned>(blocks), threads, smem, st>>>(p);
  PSK_CUDA_OK(cudaGetLastError());
  constexpr bool LF = (FLUX == PSK_FLUX_LAX_FRIEDRICHS);
  if (LF || p.bc.bc == PSK_BC_PERIODIC || p.bc.bc == PSK_BC_NEUMANN) {
    adjoint_boundary_kernel<LF><<<batch, LF ? 128 : 32, 0, st>>>(p);
    PSK_CUDA_OK(cudaGetLastError());
  }
  return PSK_OK;
}

template <int EQ, int FLUX>
int adjoint_rec(int rec, const AdjParams &p, int batch, cudaStream_t st) {
  switch (rec) {
    case PSK_REC_CONSTANT: return launch_adjoint<EQ, FLUX, PSK_REC_CONSTANT>(p, batch, st);
    case PSK_REC_WENOJS32: return launch_adjoint<EQ, FLUX, PSK_REC_WENOJS32>(p, batch, st);
    case PSK_REC_ESWENO32: return launch_adjoint<EQ, FLUX, PSK_REC_ESWENO32>(p, batch, st);
    default: return launch_adjoint<EQ, FLUX, PSK_REC_WENOJS53>(p, batch, st);
  }
}

// defined in psk_forward.cu
int launch_max_abs_public(const psk_desc *d, const double *u, int mode, double *out, cudaStream_t st);

static int run_adjoint(const psk_desc *d, const double *x, const double *v, const double *dt,
                       int64_t dt_stride, double c_v, double c_g, const double *acc, double c_acc,
                       const double *acc2, double c_acc2, double *work, double *out,
                       cudaStream_t st) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (x == nullptr || v == nullptr || out == nullptr || work == nullptr) return PSK_E_INVALID;
  if (out == x || out == v) return PSK_E_INVALID;  // tiles read their neighbours' cells
  // the ESWENO32 scheme is the upwind flux of the ESWENO32 reconstruction plus a dissipative flux of its weights
  if (d->flux == PSK_FLUX_ESWENO && (d->rec != PSK_REC_ESWENO32 || d->equation != PSK_EQ_BURGERS)) return PSK_E_INVALID;
  AdjParams p{};
  p.x = x;
  p.v = v;
  p.acc = (acc != nullptr && c_acc != 0.0) ? acc : nullptr;
  p.acc2 = (acc2 != nullptr && c_acc2 != 0.0) ? acc2 : nullptr;
  p.out = out;
  p.dt = dt;
  p.dt_stride = dt_stride;
  p.c_v = c_v;
  p.c_g = c_g;
  p.c_acc = c_acc;
  p.c_acc2 = c_acc2;
  p.speed = work;
  p.ga = work + d->batch;
  p.amax = nullptr;
  p.gspill = work + 2 * static_cast<int64_t>(d->batch);
  p.nu = d->nu;
  p.vel = d->velocity;
  p.vel_l = d->vel_l;
  p.vel_r = d->vel_r;
  p.bc = make_bc_view(d);
  p.ld = d->ld;
  p.invdx = 1.0 / d->dx;
  p.eps = d->eps;
  p.delta = d->delta;
  const bool lf = d->equation == PSK_EQ_BURGERS && d->flux == PSK_FLUX_LAX_FRIEDRICHS;
  if (lf) {
    rc = launch_max_abs_public(d, x, 2, p.speed, st);
    if (rc != PSK_OK) return rc;
    PSK_CUDA_OK(cudaMemsetAsync(p.ga, 0, sizeof(double) * d->batch, st));
    // (count, index) of the row's arg-max cells, behind the ghost-cell cotangents
    p.amax = reinterpret_cast<unsigned *>(work + (2 + 2 * static_cast<int64_t>(d->g)) * d->batch);
    PSK_CUDA_OK(cudaMemsetAsync(p.amax, 0, sizeof(double) * d->batch, st));
  }
  const int b = d->batch;
  if (d->equation == PSK_EQ_ADVECTION)
    return adjoint_rec<PSK_EQ_ADVECTION, PSK_FLUX_UPWIND>(d->rec, p, b, st);
  if (d->equation == PSK_EQ_CONTINUITY)
    return adjoint_rec<PSK_EQ_CONTINUITY, PSK_FLUX_UPWIND>(d->rec, p, b, st);
  switch (d->flux) {
    case PSK_FLUX_RUSANOV: return adjoint_rec<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV>(d->rec, p, b, st);
    case PSK_FLUX_LAX_FRIEDRICHS:
      return adjoint_rec<PSK_EQ_BURGERS, PSK_FLUX_LAX_FRIEDRICHS>(d->rec, p, b, st);
    case PSK_FLUX_UPWIND: return adjoint_rec<PSK_EQ_BURGERS, PSK_FLUX_UPWIND>(d->rec, p, b, st);
    case PSK_FLUX_ESWENO: return launch_adjoint<PSK_EQ_BURGERS, PSK_FLUX_ESWENO, PSK_REC_ESWENO32>(p, b, st);
    default: return adjoint_rec<PSK_EQ_BURGERS, PSK_FLUX_ENGQUIST_OSHER>(d->rec, p, b, st);
  }
}

}  // namespace psk

using namespace psk;

extern "C" {

/* A/B switch for the adjoint stage (see g_adjoint_variant) */
int psk_set_adjoint_variant(int variant) {
  if (variant < 0 || variant > 6) return PSK_E_INVALID;
  g_adjoint_variant = variant;
  return PSK_OK;
}

int psk_apply_operator_vjp(const psk_desc *d, const double *u, const double *v, double *out,
                           double *work, psk_stream_t stream) {
  return run_adjoint(d, u, v, nullptr, 0, 0.0, 1.0, nullptr, 0.0, nullptr, 0.0, work, out,
                     static_cast<cudaStream_t>(stream));
}

int psk_ssprk33_stage_adjoint(const psk_desc *d, const double *x, const double *v, const double *dt,
                              int64_t dt_stride, double c_v, const double *acc, double c_acc,
                              const double *acc2, double c_acc2, double *work, double *out,
                              psk_stream_t stream) {
  if (dt == nullptr) return PSK_E_INVALID;
  return run_adjoint(d, x, v, dt, dt_stride, c_v, c_v, acc, c_acc, acc2, c_acc2, work, out,
                     static_cast<cudaStream_t>(stream));
}

/* The whole reverse sweep of adjoint_step (timestepping.py:198-209) in ONE call: per step m = nsteps - 1 .. 0 the
 * stage values of the checkpointed state are recomputed (2 launches) and the three adjoint stages applied (3
 * launches), then the boundary condition of the adjoint variable (1 launch) -- all enqueued back to back from
 * here, so a small grid pays kernel launches, not Python round trips. */
int psk_ssprk33_adjoint_sweep(const psk_desc *d, const double *tape, int64_t tape_stride, int nsteps,
                              const double *dt_table, const double *ghost_table, const psk_desc *pbc, double *p,
                              double *states, double *work, double *lf_work, double *p_hist, psk_stream_t stream) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (tape == nullptr || dt_table == nullptr || p == nullptr || states == nullptr || work == nullptr || nsteps <= 0)
    return PSK_E_INVALID;
  if (pbc != nullptr) {
    rc = check_desc(pbc);
    if (rc != PSK_OK) return rc;
    if (pbc->n != d->n || pbc->g != d->g || pbc->batch != d->batch || pbc->ld != d->ld) return PSK_E_INVALID;
  }
  const int64_t state = static_cast<int64_t>(d->batch) * d->ld;
  double *k1 = states, *k2 = states + state, *lam2 = states + 2 * state, *lam1 = states + 3 * state,
         *pn = states + 4 * state;
  const int64_t ghost_block = d->ghost_ld != 0 ? static_cast<int64_t>(d->batch) * d->ghost_ld : 2 * d->g;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double *cur = p, *nxt = pn;
  for (int m = nsteps - 1; m >= 0; --m) {
    const double *u = tape + static_cast<int64_t>(m) * tape_stride;
    const double *dt = dt_table + m;
    psk_desc ds[3] = {*d, *d, *d};  // boundary data at t, t + dt, t + dt / 2
    if (ghost_table != nullptr)
      for (int s = 0; s < 3; ++s) ds[s].ghost = ghost_table + (static_cast<int64_t>(3) * m + s) * ghost_block;
    // k1, k2 of the checkpointed state: one launch where the whole-step kernel exists, else two stage launches
    // (the stored ghost cells of k1, k2 are never read: every stage kernel applies the boundary condition itself)
    rc = PSK_E_UNSUPPORTED;
    if ((d->bc == PSK_BC_DIRICHLET || d->bc == PSK_BC_NEUMANN) && ghost_table != nullptr)
      rc = psk_ssprk33_step_bc(d, u, nullptr, dt, 0, ds[0].ghost, nullptr, nullptr, k1, k2, stream);
    else if (d->bc == PSK_BC_PERIODIC)
      rc = psk_ssprk33_step_stages(d, u, k1, k2, nullptr, dt, 0, stream);
    if (rc == PSK_E_UNSUPPORTED) {
      rc = psk_ssprk33_stage(&ds[0], 1, u, u, k1, dt, 0, nullptr, lf_work, nullptr, 0, stream);
      if (rc != PSK_OK) return rc;
      rc = psk_ssprk33_stage(&ds[1], 2, u, k1, k2, dt, 0, nullptr, lf_work, nullptr, 0, stream);
    }
    if (rc != PSK_OK) return rc;
    rc = run_adjoint(&ds[2], k2, cur, dt, 0, 2.0 / 3.0, 2.0 / 3.0, nullptr, 0.0, nullptr, 0.0, work, lam2, st);
    if (rc != PSK_OK) return rc;
    rc = run_adjoint(&ds[1], k1, lam2, dt, 0, 0.25, 0.25, nullptr, 0.0, nullptr, 0.0, work, lam1, st);
    if (rc != PSK_OK) return rc;
    rc = run_adjoint(&ds[0], u, lam1, dt, 0, 1.0, 1.0, cur, 1.0 / 3.0, lam2, 0.75, work, nxt, st);
    if (rc != PSK_OK) return rc;
    if (pbc != nullptr) {  // p = apply_boundary(t, u, p) (timestepping.py:208-209)
      rc = psk_apply_boundary(pbc, nxt, cur, stream);
      if (rc != PSK_OK) return rc;
    } else {
      double *tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    if (p_hist != nullptr)
      PSK_CUDA_OK(cudaMemcpyAsync(p_hist + static_cast<int64_t>(m) * state, cur, sizeof(double) * state,
                                  cudaMemcpyDeviceToDevice, st));
  }
  if (cur != p) PSK_CUDA_OK(cudaMemcpyAsync(p, cur, sizeof(double) * state, cudaMemcpyDeviceToDevice, st));
  return PSK_OK;
}

}  // extern "C"
