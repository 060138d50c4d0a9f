// psk_adjoint_kernels.cuh -- the lean adjoint stage kernel of the hot configuration (Burgers, Rusanov,
// WENO-JS5, nu = 1, aligned rows) and its parameter block.  Device code only (no launches, no CUDA
// runtime calls), so that tests/host/ can compile this very file for the HOST with the warp
// emulation of tests/host/emu/cuda_runtime.h (tests/test_adjoint_kernel_host.py) and check the
// transposed stencil, the lane exchange and the ghost-cell spill against the reverse-mode
// derivative of the reference arithmetic without a GPU.
#pragma once

#include "psk_common.cuh"
#include "psk_math.cuh"
#include "psk_adjoint_math.cuh"

namespace psk {

struct AdjParams {
  const double *x;
  const double *v;
  const double *acc;
  const double *acc2;
  double *out;
  const double *dt;  // nullptr -> 1
  int64_t dt_stride;
  double c_v, c_g, c_acc, c_acc2;
  double *speed;   // [batch] global LF speed (input)
  double *ga;      // [batch] cotangent of the LF speed (accumulated here)
  unsigned *amax;  // [batch][2] Lax-Friedrichs, lean kernel: how many cells of the row hold the speed max |w| and the
                   // array index of one of them (adjoint_boundary_kernel scans the row only when there are several)
  double *gspill;  // [batch][2g] cotangents that landed on ghost cells
  const double *nu;
  const double *vel;
  const double *vel_l;
  const double *vel_r;
  BcView bc;
  int64_t ld;
  double invdx, eps;
  double delta;  // ESWENO32 scheme only (burgers/schemes.py:243)
  int tiles_per_row;
  int prescaled;  // gspill already carries the factor c_g dt (lean kernel)
};

// ---------------------------------------------------------------------------
// Lean warp form of the adjoint stage for the hot configuration (Burgers, Rusanov, WENO-JS5,
// nu = 1, aligned rows) -- the transposed counterpart of stage_warp_fast_kernel.  Same data
// layout as adjoint_warp_kernel (one warp = 128 consecutive cells, lanes 0 and 31 are halo
// lanes, everything between lanes travels by shuffle), different arithmetic and schedule:
//   * psk_adjoint_math.cuh: the derivative of the PRODUCT form of the weights (no 1 / e_k),
//     accumulated on the seven first differences a lane touches; the conversion to cell
//     cotangents is one subtraction per cell at the end;
//   * every cotangent is pre-scaled by c_g dt / dx, so the epilogue is lin + gi;
//   * all global loads (x, v, acc, acc2) are issued at the top and lin is formed at once;
//   * the forward pass of a cell (Weno5State) is kept and reused by its VJP instead of being
//     recomputed: cells are processed in the order fwd(3), fwd(0) | exchange | face(0), fwd(1),
//     face(1), vjp(0), fwd(2), face(2), vjp(1), face(3), vjp(2), face(4), vjp(3), so at most three
//     states are alive;
//   * the Rusanov speed max(|w_j|, |w_j+1|), its arg-max and the signs are integer work on the
//     bit patterns (ALU pipe), not FP64 compares;
//   * CTAs of 4 warps: 4-5 CTAs per SM drift apart, so the load phase of one hides behind the
//     arithmetic of the others.
struct LeanFace {
  double gR;  // cotangent of the right value of the left cell (j)
  double gL;  // cotangent of the left value of the right cell (p)
  double dj, dp;  // direct terms on w_j, w_p (through the speed)
};

// +-1, +-1/2 or 0 as a double: sign(w) times `half_exp` (0x3FF00000 -> 1, 0x3FE00000 -> 1/2, 0 -> 0)
__device__ __forceinline__ double signed_unit(double w, unsigned hi_bits) {
  const unsigned sign = static_cast<unsigned>(__double2hiint(w)) & 0x80000000u;
  return __hiloint2double(static_cast<int>(sign | hi_bits), 0);
}

// hG = (1/2) c_g dt (v_p - v_j) / dx ;  Phi = 1/4 (urj^2 + ulp^2) - 1/2 a (ulp - urj)
__device__ __forceinline__ LeanFace lean_face(double hG, double urj, double ulp, double wj, double wp) {
  const unsigned long long bj = static_cast<unsigned long long>(__double_as_longlong(wj)) & 0x7fffffffffffffffull;
  const unsigned long long bp = static_cast<unsigned long long>(__double_as_longlong(wp)) & 0x7fffffffffffffffull;
  const double a = __longlong_as_double(static_cast<long long>(bj > bp ? bj : bp));
  LeanFace o;
  o.gR = hG * (urj + a);
  o.gL = hG * (ulp - a);
  const double da = hG * (urj - ulp);
  // jnp.maximum: the larger argument takes the gradient, ties split 1/2 - 1/2; abs'(0) = 0
  const unsigned ej = bj > bp ? 0x3FF00000u : ((bj == bp && bj != 0ull) ? 0x3FE00000u : 0u);
  const unsigned ep = bp > bj ? 0x3FF00000u : ((bj == bp && bp != 0ull) ? 0x3FE00000u : 0u);
  o.dj = da * signed_unit(wj, ej);
  o.dp = da * signed_unit(wp, ep);
  return o;
}

// The same face with the viscosity nu of the face (scalar.py:231-234: the speed is multiplied by grid.df ** (alpha - 1))
__device__ __forceinline__ LeanFace lean_face_nu(double hG, double urj, double ulp, double wj, double wp, double nu) {
  const unsigned long long bj = static_cast<unsigned long long>(__double_as_longlong(wj)) & 0x7fffffffffffffffull;
  const unsigned long long bp = static_cast<unsigned long long>(__double_as_longlong(wp)) & 0x7fffffffffffffffull;
  const double a = nu * __longlong_as_double(static_cast<long long>(bj > bp ? bj : bp));
  LeanFace o;
  o.gR = hG * (urj + a);
  o.gL = hG * (ulp - a);
  const double da = (hG * nu) * (urj - ulp);
  const unsigned ej = bj > bp ? 0x3FF00000u : ((bj == bp && bj != 0ull) ? 0x3FE00000u : 0u);
  const unsigned ep = bp > bj ? 0x3FF00000u : ((bj == bp && bp != 0ull) ? 0x3FE00000u : 0u);
  o.dj = da * signed_unit(wj, ej);
  o.dp = da * signed_unit(wp, ep);
  return o;
}

// Global Lax-Friedrichs flux (scalar.py:258-278): the speed `a` = nu max |w| over the whole row is the same number at
// every face, so a face has no direct terms on its two cells; `dj` carries the cotangent of the row's speed instead
// (hG nu (urj - ulp): summed over the faces of the row into AdjParams::ga, distributed to the arg-max cells by
// adjoint_boundary_kernel<true>), `dp` is zero
__device__ __forceinline__ LeanFace lean_face_lf(double hG, double urj, double ulp, double speed, double nu) {
  const double a = nu * speed;
  LeanFace o;
  o.gR = hG * (urj + a);
  o.gL = hG * (ulp - a);
  o.dj = (hG * nu) * (urj - ulp);
  o.dp = 0.0;
  return o;
}

// own cells of a lane: state, cotangent, linear part of the result; plus the one cell the
// two edge lanes need beyond the shuffled halo (lane 0: cell c0 - 1, lane 31: cell c0 + R)
struct LeanIn {
  double w[4], v[4], lin[4];
  double extra;
};

__device__ __forceinline__ void lean_lin(const AdjParams &p, LeanIn &in, const double2 &r0, const double2 &r1,
                                         const double2 &a0, const double2 &a1, const double2 &b0,
                                         const double2 &b1) {
  in.v[0] = r0.x; in.v[1] = r0.y; in.v[2] = r1.x; in.v[3] = r1.y;
  in.lin[0] = fma(p.c_acc2, b0.x, fma(p.c_acc, a0.x, p.c_v * r0.x));
  in.lin[1] = fma(p.c_acc2, b0.y, fma(p.c_acc, a0.y, p.c_v * r0.y));
  in.lin[2] = fma(p.c_acc2, b1.x, fma(p.c_acc, a1.x, p.c_v * r1.x));
  in.lin[3] = fma(p.c_acc2, b1.y, fma(p.c_acc, a1.y, p.c_v * r1.y));
}

// synchronous loads (all issued before the first use)
__device__ __forceinline__ void lean_load(const AdjParams &p, int row, int c0, int lane, bool inside,
                                          LeanIn &in) {
  constexpr int R = 4;
  const int g = p.bc.g, nx = p.bc.nx;
  const int64_t base = static_cast<int64_t>(row) * p.ld;
  const int64_t off = base + g + c0;
  const double *__restrict__ xrow = p.x + base;
  if (inside) {
    const double2 q0 = *reinterpret_cast<const double2 *>(p.x + off);
    const double2 q1 = *reinterpret_cast<const double2 *>(p.x + off + 2);
    const double2 r0 = *reinterpret_cast<const double2 *>(p.v + off);
    const double2 r1 = *reinterpret_cast<const double2 *>(p.v + off + 2);
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0;
    if (p.acc != nullptr) {
      a0 = *reinterpret_cast<const double2 *>(p.acc + off);
      a1 = *reinterpret_cast<const double2 *>(p.acc + off + 2);
    }
    if (p.acc2 != nullptr) {
      b0 = *reinterpret_cast<const double2 *>(p.acc2 + off);
      b1 = *reinterpret_cast<const double2 *>(p.acc2 + off + 2);
    }
    in.w[0] = q0.x; in.w[1] = q0.y; in.w[2] = q1.x; in.w[3] = q1.y;
    lean_lin(p, in, r0, r1, a0, a1, b0, b1);
  } else {
    const double *__restrict__ vrow = p.v + base;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = g + c0 + r;
      const bool in_row = (i >= 0 && i < nx);
      in.w[r] = load_w(p.bc, xrow, row, i);
      in.v[r] = in_row ? vrow[i] : 0.0;
      double l = p.c_v * in.v[r];
      if (in_row && p.acc != nullptr) l = fma(p.c_acc, p.acc[base + i], l);
      if (in_row && p.acc2 != nullptr) l = fma(p.c_acc2, p.acc2[base + i], l);
      in.lin[r] = l;
    }
  }
  in.extra = 0.0;
  if (lane == 0) in.extra = load_w(p.bc, xrow, row, g + c0 - 1);
  if (lane == 31) in.extra = load_w(p.bc, xrow, row, g + c0 + R);
}

// everything after the loads: window exchange, forward states, faces, VJPs, spill exchange, store
// slin: where the linear part was parked (shared memory, stride 128 doubles) or nullptr (in.lin);
// RECOMP3: recompute the forward state of cell 3 before its faces instead of holding it since
// the exchange at the top (39 more FP64 instructions per lane, 26 registers fewer in between)
// FLUX: PSK_FLUX_RUSANOV or PSK_FLUX_LAX_FRIEDRICHS; NU: AdjParams::nu holds the viscosity of every face (alpha != 1)
template <bool RECOMP3, int FLUX = PSK_FLUX_RUSANOV, bool NU = false>
__device__ __forceinline__ void lean_compute_store(const AdjParams &p, int row, int c0, int lane, bool inside,
                                                   const LeanIn &in, const double *slin) {
  constexpr int R = 4;
  constexpr unsigned kFull = 0xffffffffu;
  const int g = p.bc.g, n = p.bc.n, nx = p.bc.nx;
  const int64_t base = static_cast<int64_t>(row) * p.ld;
  const int64_t off = base + g + c0;
  const double cgdt = p.c_g * (p.dt != nullptr ? p.dt[static_cast<int64_t>(row) * p.dt_stride] : 1.0);

  double w[R + 6], vc[R + 2];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    w[3 + r] = in.w[r];
    vc[1 + r] = in.v[r];
  }
  w[1] = __shfl_up_sync(kFull, w[5], 1);
  w[2] = __shfl_up_sync(kFull, w[6], 1);
  w[7] = __shfl_down_sync(kFull, w[3], 1);
  w[8] = __shfl_down_sync(kFull, w[4], 1);
  if (lane == 0) w[2] = in.extra;
  if (lane == 31) w[7] = in.extra;
  vc[0] = __shfl_up_sync(kFull, vc[R], 1);
  vc[R + 1] = __shfl_down_sync(kFull, vc[1], 1);

  // ---- (1/2) c_g dt (v_k - v_{k-1}) / dx for the faces of the lane; zero at the array ends
  const double hs = 0.5 * cgdt * p.invdx;
  double hG[R + 1], nuf[R + 1];
#pragma unroll
  for (int f = 0; f <= R; ++f) {
    const int k = g + c0 + f;  // array index of the face: between cells k - 1 and k
    const bool valid = (k >= 1 && k <= nx - 1);
    hG[f] = valid ? (vc[f + 1] - vc[f]) * hs : 0.0;
    nuf[f] = (NU && valid) ? p.nu[k - 1] : 1.0;
  }
  constexpr bool kLF = FLUX == PSK_FLUX_LAX_FRIEDRICHS;
  const double speed = kLF ? p.speed[row] : 0.0;
  // the face between the window cells (f + 2 | f + 3)
  auto face = [&](int f, double urj, double ulp, double wj, double wp) -> LeanFace {
    if (kLF) return lean_face_lf(hG[f], urj, ulp, speed, nuf[f]);
    if (NU) return lean_face_nu(hG[f], urj, ulp, wj, wp, nuf[f]);
    return lean_face(hG[f], urj, ulp, wj, wp);
  };

  // ---- first differences in sixths (t[k]: cells k, k+1 of the window) and (13/3) dd^2 + eps/9
  const double eps9 = p.eps * (1.0 / 9.0);
  double t[R + 4], pq[R + 3];  // used: t[1..7], pq[1..6]
#pragma unroll
  for (int k = 1; k <= R + 3; ++k) t[k] = (1.0 / 6.0) * (w[k + 1] - w[k]);
#pragma unroll
  for (int k = 1; k <= R + 2; ++k) {
    const double dd = t[k + 1] - t[k];
    pq[k] = fma((13.0 / 3.0) * dd, dd, eps9);
  }
  // cell r (window index r + 3) uses t[r+1 .. r+4], pq[r+1 .. r+3]; its cotangents go to Tk[r .. r+3]
  double Tk[R + 3];
#pragma unroll
  for (int k = 0; k < R + 3; ++k) Tk[k] = 0.0;
  double o[R];

  const Weno5State F3 = weno53_state(t[4], t[5], t[6], t[7], pq[4], pq[5], pq[6]);
  const Weno5State F0 = weno53_state(t[1], t[2], t[3], t[4], pq[1], pq[2], pq[3]);
  const double ur3 = w[6] + F3.uR, ul0 = w[3] + F0.uL;
  const double ur_left = __shfl_up_sync(kFull, ur3, 1);
  const double ul_right = __shfl_down_sync(kFull, ul0, 1);

  const LeanFace f0 = face(0, ur_left, ul0, w[2], w[3]);
  const Weno5State F1 = weno53_state(t[2], t[3], t[4], t[5], pq[2], pq[3], pq[4]);
  const LeanFace f1 = face(1, w[3] + F0.uR, w[4] + F1.uL, w[3], w[4]);
  o[0] = kLF ? (f1.gR + f0.gL) : (f0.dp + f1.dj) + (f1.gR + f0.gL);
  weno53_vjp_acc(F0, t[1], t[2], t[3], t[4], f1.gR, f0.gL, Tk[0], Tk[1], Tk[2], Tk[3]);

  const Weno5State F2 = weno53_state(t[3], t[4], t[5], t[6], pq[3], pq[4], pq[5]);
  const LeanFace f2 = face(2, w[4] + F1.uR, w[5] + F2.uL, w[4], w[5]);
  o[1] = kLF ? (f2.gR + f1.gL) : (f1.dp + f2.dj) + (f2.gR + f1.gL);
  weno53_vjp_acc(F1, t[2], t[3], t[4], t[5], f2.gR, f1.gL, Tk[1], Tk[2], Tk[3], Tk[4]);

  double t7b = t[7];
#ifndef PSK_HOST_EMU
  if (RECOMP3) asm volatile("" : "+d"(t7b));  // keeps the compiler from merging the two evaluations
#endif
  const Weno5State F3b = RECOMP3 ? weno53_state(t[4], t[5], t[6], t7b, pq[4], pq[5], pq[6]) : F3;
  const LeanFace f3 = face(3, w[5] + F2.uR, w[6] + F3b.uL, w[5], w[6]);
  o[2] = kLF ? (f3.gR + f2.gL) : (f2.dp + f3.dj) + (f3.gR + f2.gL);
  weno53_vjp_acc(F2, t[3], t[4], t[5], t[6], f3.gR, f2.gL, Tk[2], Tk[3], Tk[4], Tk[5]);

  const LeanFace f4 = face(4, RECOMP3 ? w[6] + F3b.uR : ur3, ul_right, w[6], w[7]);
  o[3] = kLF ? (f4.gR + f3.gL) : (f3.dp + f4.dj) + (f4.gR + f3.gL);
  weno53_vjp_acc(F3b, t[4], t[5], t[6], t7b, f4.gR, f3.gL, Tk[3], Tk[4], Tk[5], Tk[6]);

  if (kLF && p.amax != nullptr && lane >= 1 && lane <= 30) {
    // the arg-max cells of the row among the cells this lane stores (ghost cells included, like jnp.max over the array)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = g + c0 + r;
      if (i >= 0 && i < nx && fabs(w[3 + r]) == speed) {
        atomicAdd(p.amax + 2 * row, 1u);
        atomicMax(p.amax + 2 * row + 1, static_cast<unsigned>(i));
      }
    }
  }
  if (kLF) {
    // cotangent of the row's speed: every face of the row once (faces 1..4 of the lanes that store; face 0 is the
    // previous lane's face 4), already scaled by c_g dt like everything here (AdjParams::prescaled)
    double ga_part = (lane >= 1 && lane <= 30) ? ((f1.dj + f2.dj) + (f3.dj + f4.dj)) : 0.0;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) ga_part += __shfl_xor_sync(kFull, ga_part, o2);
    if (lane == 0) atomicAdd(p.ga + row, ga_part);
  }
  // ---- contributions of the neighbour lanes' cells to the first differences around my cells
  const double fl5 = __shfl_up_sync(kFull, Tk[5], 1);
  const double fl6 = __shfl_up_sync(kFull, Tk[6], 1);
  const double fr0 = __shfl_down_sync(kFull, Tk[0], 1);
  const double fr1 = __shfl_down_sync(kFull, Tk[1], 1);
  Tk[1] += fl5;
  Tk[2] += fl6;
  Tk[4] += fr0;
  Tk[5] += fr1;

  if (lane >= 1 && lane <= 30) {
    double gi[R], lin[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      gi[r] = fma(1.0 / 6.0, Tk[r + 1] - Tk[r + 2], o[r]);
      lin[r] = (slin != nullptr) ? slin[128 * r] : in.lin[r];
    }
    if (inside) {
      // inside => interior cells only (no ghost among them)
      *reinterpret_cast<double2 *>(p.out + off) = make_double2(lin[0] + gi[0], lin[1] + gi[1]);
      *reinterpret_cast<double2 *>(p.out + off + 2) = make_double2(lin[2] + gi[2], lin[3] + gi[3]);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = g + c0 + r;
        if (i < 0 || i >= nx) continue;
        const bool ghost = (p.bc.bc != PSK_BC_NONE) && (i < g || i >= nx - g);
        if (ghost) {
          // ghost cells of x do not influence L; what landed on them goes back through the
          // transpose of apply_boundary (already scaled by c_g dt: p.prescaled)
          p.out[base + i] = lin[r];
          p.gspill[static_cast<int64_t>(row) * 2 * g + (i < g ? i : i - n)] = gi[r];
        } else {
          p.out[base + i] = lin[r] + gi[r];
        }
      }
    }
  }
}

// VAR bit 0: park the linear part in shared memory; bit 1: recompute the state of cell 3
template <int MINB, int VAR, int FLUX = PSK_FLUX_RUSANOV, bool NU = false>
__global__ void __launch_bounds__(128, MINB)
adjoint_lean_kernel(const AdjParams p, int chunks_per_row) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31;
  const int chunk = static_cast<int>(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) - 1;
  if (chunk >= chunks_per_row) return;
  const int row = blockIdx.y + blockIdx.z * gridDim.y;
  const int c0 = chunk * 30 * R - R + R * lane;  // interior coordinates; array index = g + c0
  const bool inside = (c0 >= 0) && (c0 + R <= p.bc.n);
  LeanIn in;
  lean_load(p, row, c0, lane, inside, in);
  if (VAR & 1) {
    __shared__ double slin[4][128];
#pragma unroll
    for (int r = 0; r < 4; ++r) slin[r][threadIdx.x] = in.lin[r];
    lean_compute_store<(VAR & 2) != 0, FLUX, NU>(p, row, c0, lane, inside, in, &slin[0][threadIdx.x]);
  } else {
    lean_compute_store<(VAR & 2) != 0, FLUX, NU>(p, row, c0, lane, inside, in, nullptr);
  }
}

}  // namespace psk
