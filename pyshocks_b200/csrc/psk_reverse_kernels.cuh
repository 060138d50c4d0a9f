// psk_reverse_kernels.cuh -- ONE launch per reverse SSPRK33 step (Burgers, Rusanov, WENO-JS5, FAST
// math, nu = 1, periodic aligned rows: BASELINE config 5): what the reference obtains per step as
// jax.jacfwd(advance)(dt, t, u).T @ p (timestepping.py:174, :198-209), here
//
//     k1 = u + dt L(u),  k2 = 3/4 u + 1/4 (k1 + dt L(k1))                  (recomputed, timestepping.py:314-317)
//     lam2 = 2/3 (p' + dt J(k2)^T p'),  lam1 = 1/4 (lam2 + dt J(k1)^T lam2)
//     p = 1/3 p' + 3/4 lam2 + lam1 + dt J(u)^T lam1                        
//
// by temporal blocking over the five stages: u and p' are read once, p is written once (24 B per
// cell-step instead of the 144 B of five streamed launches), k1, k2, lam2, lam1 never leave the SM.
//
//   * a WARP owns a window of 32 C consecutive cells of one row, lane l the run C l .. C l + C - 1
//     (C = 12 / 16 / 20 / 24).  The row is a periodic ring: a window that reaches over a row end holds
//     the periodic images, every face is a regular face, nothing lands on ghost cells;
//   * the arrays of a window live in LANE-PRIVATE slots of shared memory (pair k of lane l at
//     [k][l]: conflict-free 128-bit accesses, no barrier of any kind -- a lane only ever touches its
//     own slots; CTA = one warp), two halo cells per side per lane filled by shuffles after the
//     stage that produced the array.  THREE arrays suffice (17 KB per warp at C = 16, 12 warps per
//     SM): u | k1 | k2, then p' over u, lam2 over k2 and 1/3 p' + 3/4 lam2 over p', lam1 over k1, u
//     once more (an L2 hit) over lam2, p over the accumulator;
//   * a stage is a STREAM along the lane's run, four cells per loop iteration, with the arithmetic
//     of the stage kernels (psk_fast_kernels.cuh: same expressions, bit-identical k1, k2;
//     psk_adjoint_math.cuh / lean_face for the adjoint): the forward stream lags its flux difference
//     one cell behind the reconstruction, the adjoint stream keeps ONE forward state alive, finishes a
//     cell two cells behind its vector-Jacobian product (when the last cotangent of its first
//     differences has arrived) and overwrites the stage array it differentiates in place (lam2 over
//     k2, lam1 over k1).  What crosses lanes is exchanged once per stage: the face values at the run
//     ends before / after the stream, the four first-difference cotangents that belong to the
//     neighbours' cells after it;
//   * every stage invalidates cells at the two ends of the WINDOW (lanes 0 and 31 shuffle with
//     themselves; u and p' of the halo cells are real): k1 is wrong in 1 cell per side, k2 in 4, lam2
//     in 9, lam1 in 12, p in 15, so the window cells [16, 32 C - 16) are stored and consecutive
//     windows start 32 C - 32 cells apart (128-byte aligned runs for C = 16).
//
// Device code only (no launches): tests/host/reverse_kernel_host.cpp compiles this very file for
// the HOST under the warp emulation and checks it against reverse-mode differentiation of the
// reference arithmetic and against the five-launch path.
#pragma once

#include "psk_common.cuh"
#include "psk_math.cuh"
#include "psk_adjoint_math.cuh"
#include "psk_fast_kernels.cuh"
#include "psk_adjoint_kernels.cuh"

namespace psk {

struct RevParams {
  const double *u;    // checkpointed state u^m                 [batch][ld]
  const double *pin;  // cotangent p^{m+1}                      [batch][ld]
  double *pout;       // cotangent p^m (interior cells only)    [batch][ld]
  const double *dt;
  int64_t dt_stride;
  int64_t ld;
  double invdx, eps;
  int n, g;
  int tiles_per_row;
  int c_last;               // cells per lane of the LAST window of a row (8 <= c_last <= C, multiple of 4)
  double *dbg_k1, *dbg_k2;  // optional: the recomputed stage values of the stored cells (or nullptr)
  int bc_none;              // 1: a slab of a larger grid -- window cells beyond the row ends are the row's STORED ghost
                            // cells (g >= 16, filled by the neighbouring slabs) instead of the periodic images
  // Dirichlet rows (reverse_step_kernel<.., DIR = true>; scalar.py:418-427): the three ghost cells per side of the
  // stage inputs u, k1, k2 hold the data of the stage times t, t + dt, t + dt / 2 (timestepping.py:314-319), block
  // s of `ghost3` at ghost3 + s * ghost_block, row r at + r * ghost_ld, left ghost cells first (psk_ssprk33_step_bc)
  const double *ghost3;
  int64_t ghost_ld, ghost_block;
};

constexpr int kRevHalo = 16;   // invalid window cells per side (15 needed)
constexpr int kRevScratch = 6;  // parked doubles per lane (first two cells of an adjoint stream)

template <int C>
struct RevGeometry {
  static_assert(C % 4 == 0 && C >= 8, "runs are streamed four cells at a time");
  static constexpr int kWindow = 32 * C;
  static constexpr int kEmit = kWindow - 2 * kRevHalo;
  static constexpr int kSlots = C + 4;               // run cells -2 .. C + 1
  static constexpr int kPairs = kSlots / 2;
  static constexpr int kArrayDoubles = kSlots * 32;  // one array of one warp
  static constexpr int kSmemDoubles = 3 * kArrayDoubles + kRevScratch * 32;
};

// Windows of a row for run length CM: all but the last hold 32 CM cells and store 32 CM - 32; the last one is as
// short as the cells left over allow (runs of c_last cells), so a row of n = 8192 cells costs 13 x 640 + 384 cells
// of work at CM = 20 instead of 14 x 640.
inline void rev_tiling(int n, int CM, int &tiles, int &c_last) {
  const int emit = 32 * CM - 2 * kRevHalo;
  tiles = (n + emit - 1) / emit;
  const int rem = n - (tiles - 1) * emit;
  c_last = 4 * ((rem + 2 * kRevHalo + 127) / 128);
  if (c_last < 8) c_last = 8;
  if (c_last > CM) c_last = CM;
}
inline long long rev_work(int n, int CM) {
  int tiles, c_last;
  rev_tiling(n, CM, tiles, c_last);
  return static_cast<long long>(tiles - 1) * 32 * CM + 32 * c_last;
}

// lane-private view of one window array: run cell j (-2 <= j <= C + 1) of this lane
struct LaneArray {
  double2 *p;  // &array[lane] as pairs; pair k of the lane at p[32 k]
  __device__ __forceinline__ double2 ld2(int j) const { return p[((j + 2) >> 1) * 32]; }  // j even
  __device__ __forceinline__ void st2(int j, double a, double b) const { p[((j + 2) >> 1) * 32] = make_double2(a, b); }
  __device__ __forceinline__ double ld1(int j) const {
    return reinterpret_cast<const double *>(p + ((j + 2) >> 1) * 32)[(j + 2) & 1];
  }
  __device__ __forceinline__ void st1(int j, double a) const {
    reinterpret_cast<double *>(p + ((j + 2) >> 1) * 32)[(j + 2) & 1] = a;
  }
};

// scaled Rusanov flux 4 F of the forward kernels (psk_fast_kernels.cuh, step_stage_rhs)
__device__ __forceinline__ double rev_flux4(double urj, double ulp, double m2j, double m2p) {
  return fma(umax_neg(m2j, m2p), ulp - urj, fma(urj, urj, ulp * ulp));
}

// stage 1: k1 = x + cdt dF; stage 2: k2 = 3/4 u0 + 1/4 (x + cdt dF) -- the expressions of the stage kernels
__device__ __forceinline__ double rev_combine(bool stage2, double x, double u0, double cdt, double dF) {
  const double k = fma(cdt, dF, x);
  return stage2 ? fma(0.25, k, 0.75 * u0) : k;
}

// ---------------------------------------------------------------------------
// forward stage on the lane's run: Y = x + cdt dF(X) (stage 1) or 3/4 U0 + 1/4 (x + cdt dF(X)) (stage 2),
// then the two halo cells per side of Y from the neighbour lanes.  X holds run cells -2 .. C + 1.
// The stage is a RUN-TIME argument (as is the mode of the adjoint stage below): both forward stages
// and all three adjoint stages execute the same instructions, which keeps the kernel's code inside the
// 32 KB instruction cache -- with one inlined copy per stage (61 KB) the warps of an SM, each at its own
// place in the code, stalled on instruction fetch more than on anything else.
__device__ __forceinline__ void rev_forward_stage(const int C, bool stage2, const LaneArray X, const LaneArray Y,
                                                  const LaneArray U0, double cdt, double eps9) {
  constexpr unsigned kFull = 0xffffffffu;
  double ur_prev = 0.0, F_prev = 0.0;  // ur of cell j0 - 1, flux of the face (j0 - 2 | j0 - 1)
  double ul_first = 0.0, F_first = 0.0;  // ul of cell 0 and the flux of the face (0 | 1), for the end of the stream
  double y_prev = 0.0;                   // result of the cell j0 - 2
  // first differences (in sixths) and second-difference terms are carried from one iteration to the next: every
  // iteration forms only the four new ones (same expressions, hence the same bits, as forming all of them)
  double tc0, tc1, tc2, pc0, pc1;
  {
    const double2 a = X.ld2(-2), b = X.ld2(0);
    tc0 = __dmul_rn(1.0 / 6.0, a.y - a.x);
    tc1 = __dmul_rn(1.0 / 6.0, b.x - a.y);
    tc2 = __dmul_rn(1.0 / 6.0, b.y - b.x);
    const double d0 = tc1 - tc0, d1 = tc2 - tc1;
    pc0 = fma((13.0 / 3.0) * d0, d0, eps9);
    pc1 = fma((13.0 / 3.0) * d1, d1, eps9);
  }
#pragma unroll 1
  for (int j0 = 0; j0 < C; j0 += 4) {
    double w[8];  // cells j0 - 2 .. j0 + 5
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double2 q = X.ld2(j0 - 2 + 2 * k);
      w[2 * k] = q.x;
      w[2 * k + 1] = q.y;
    }
    double t[7], pq[6];  // t[k]: interval (j0 - 2 + k, j0 - 1 + k); pq[k]: centred at the cell j0 - 1 + k
    t[0] = tc0; t[1] = tc1; t[2] = tc2;
    pq[0] = pc0; pq[1] = pc1;
#pragma unroll
    for (int k = 3; k < 7; ++k) t[k] = __dmul_rn(1.0 / 6.0, w[k + 1] - w[k]);
#pragma unroll
    for (int k = 2; k < 6; ++k) {
      const double dd = t[k + 1] - t[k];
      pq[k] = fma((13.0 / 3.0) * dd, dd, eps9);
    }
    tc0 = t[4]; tc1 = t[5]; tc2 = t[6];
    pc0 = pq[4]; pc1 = pq[5];
    double m2[5];  // -2 |w| of the cells j0 - 1 .. j0 + 3
#pragma unroll
    for (int k = 0; k < 5; ++k) m2[k] = -2.0 * fabs(w[k + 1]);
    double ul[4], ur[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const Weno5Pair o = weno53_pair_lean(w[m + 2], t[m], t[m + 1], t[m + 2], t[m + 3], pq[m], pq[m + 1], pq[m + 2]);
      ul[m] = o.ul;
      ur[m] = o.ur;
    }
    double F[4];  // faces (j0 - 1 + m | j0 + m)
#pragma unroll
    for (int m = 0; m < 4; ++m) F[m] = rev_flux4(m == 0 ? ur_prev : ur[m - 1], ul[m], m2[m], m2[m + 1]);
    if (j0 == 0) {
      ul_first = ul[0];
      F_first = F[1];
    }
    // cells j0 - 1 .. j0 + 2 are complete; stored as the aligned pairs (j0 - 2, j0 - 1), (j0, j0 + 1) with
    // the cell j0 - 2 carried from the previous iteration (run cells -2, -1 are halo slots, rewritten below)
    double y[4];
    const double2 ua = U0.ld2(j0 - 2), ub = U0.ld2(j0), uc = U0.ld2(j0 + 2);
    const double u0v[4] = {ua.y, ub.x, ub.y, uc.x};
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double dF = (m == 0 ? F_prev : F[m - 1]) - F[m];
      y[m] = rev_combine(stage2, w[m + 1], u0v[m], cdt, dF);
    }
    Y.st2(j0 - 2, y_prev, y[0]);
    Y.st2(j0, y[1], y[2]);
    y_prev = y[3];
    ur_prev = ur[3];
    F_prev = F[3];
  }
  // the two run ends: the face values beyond them come from the neighbour lanes
  const double ur_left = __shfl_up_sync(kFull, ur_prev, 1);
  const double ul_right = __shfl_down_sync(kFull, ul_first, 1);
  const double2 a = X.ld2(-2), b = X.ld2(0), c = X.ld2(C - 2), d = X.ld2(C);
  const double F0 = rev_flux4(ur_left, ul_first, -2.0 * fabs(a.y), -2.0 * fabs(b.x));
  const double FC = rev_flux4(ur_prev, ul_right, -2.0 * fabs(c.y), -2.0 * fabs(d.x));
  const double2 uf = U0.ld2(0), ue = U0.ld2(C - 2);
  const double y0 = rev_combine(stage2, b.x, uf.x, cdt, F0 - F_first);
  const double yl = rev_combine(stage2, c.y, ue.y, cdt, F_prev - FC);
  const double y1 = Y.ld2(0).y;  // cell 1 (cell 0 of that pair is still the provisional value)
  Y.st2(0, y0, y1);
  Y.st2(C - 2, y_prev, yl);
  Y.st2(-2, __shfl_up_sync(kFull, y_prev, 1), __shfl_up_sync(kFull, yl, 1));
  Y.st2(C, __shfl_down_sync(kFull, y0, 1), __shfl_down_sync(kFull, y1, 1));
}

// ---------------------------------------------------------------------------
// adjoint stage on the lane's run:
//     OUT = c_v V [+ A] + c_g dt J_L(X)^T V      (hs = c_g dt / (2 dx))
// X: run cells -2 .. C + 1, V: run cells -1 .. C.  OUT may be X (in place: the stream reads X at
// least two cells ahead of what it writes) or A.  Afterwards OUT's halo cells -1 and C hold the
// neighbours' values (OUT is the V of the next stage).
//   MODE 0: as above without A;  MODE 1: also A = 1/3 V + 3/4 OUT (A aliases V: the accumulator
//   1/3 p' + 3/4 lam2 of the last stage takes the place of p');  MODE 2: with the term + A.
__device__ __forceinline__ void rev_adjoint_stage(const int C, const int MODE, const LaneArray X, const LaneArray V,
                                                  const LaneArray OUT, const LaneArray A, double c_v, double hs,
                                                  double eps9, double *park) {
  constexpr unsigned kFull = 0xffffffffu;
  double tc[4], pc[3];  // carried: t of the intervals (j0-2, j0-1) .. (j0+1, j0+2), pq centred at j0-1 .. j0+1
  auto state_at = [&](const double (&w)[6]) {  // state of the cell w[2] from the cells w[0..4] (+ w[5] unused)
#pragma unroll
    for (int k = 0; k < 4; ++k) tc[k] = (1.0 / 6.0) * (w[k + 1] - w[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dd = tc[k + 1] - tc[k];
      pc[k] = fma((13.0 / 3.0) * dd, dd, eps9);
    }
    return weno53_state(tc[0], tc[1], tc[2], tc[3], pc[0], pc[1], pc[2]);
  };
  // ---- the face values beyond the run ends: ur of my last cell to the right, ul of my first cell to the left
  Weno5State S;  // state of the cell the stream is at
  double ur_left, ul_right;
  {
    double w[6];
    const double2 q0 = X.ld2(C - 4), q1 = X.ld2(C - 2), q2 = X.ld2(C);
    w[0] = q0.y; w[1] = q1.x; w[2] = q1.y; w[3] = q2.x; w[4] = q2.y; w[5] = 0.0;
    const Weno5State SL = state_at(w);
    ur_left = __shfl_up_sync(kFull, w[2] + SL.uR, 1);
  }
  LeanFace fp;  // the face to the left of the cell the stream is at
  {
    double w[6];
    const double2 q0 = X.ld2(-2), q1 = X.ld2(0), q2 = X.ld2(2);
    w[0] = q0.x; w[1] = q0.y; w[2] = q1.x; w[3] = q1.y; w[4] = q2.x; w[5] = q2.y;
    S = state_at(w);
    const double ul0 = w[2] + S.uL;
    ul_right = __shfl_down_sync(kFull, ul0, 1);
    const double hG = hs * (V.ld2(0).x - V.ld2(-2).y);
    fp = lean_face(hG, ur_left, ul0, w[1], w[2]);
  }
  double Tprev = 0.0, Tc0 = 0.0, Tc1 = 0.0, Tc2 = 0.0;  // cotangents of the intervals (j0-3,j0-2) .. (j0,j0+1)
  double o_m2 = 0.0, o_m1 = 0.0;                        // direct terms of the cells j0 - 2, j0 - 1
#pragma unroll 1
  for (int j0 = 0; j0 < C; j0 += 4) {
    const bool more = (j0 + 4 < C);
    double w[10];  // cells j0 - 2 .. j0 + 7; used: j0 .. j0 + 6 (the last two only while another iteration follows)
    w[0] = w[1] = 0.0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const double2 q = X.ld2(j0 - 2 + 2 * k);
      w[2 * k] = q.x;
      w[2 * k + 1] = q.y;
    }
    w[8] = w[9] = 0.0;
    if (more) {
      const double2 q = X.ld2(j0 + 6);
      w[8] = q.x;
      w[9] = q.y;
    }
    double v[6];  // cells j0 .. j0 + 5 (the last one unused)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double2 q = V.ld2(j0 + 2 * k);
      v[2 * k] = q.x;
      v[2 * k + 1] = q.y;
    }
    // t[k]: interval (j0 - 2 + k, j0 - 1 + k); pq[k]: centred at the cell j0 - 1 + k.  The first four / three were
    // formed by the previous iteration (or with the state of cell 0): only the new ones are computed
    double t[8], pq[7];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = tc[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) pq[k] = pc[k];
#pragma unroll
    for (int k = 4; k < 8; ++k) t[k] = (1.0 / 6.0) * (w[k + 1] - w[k]);
#pragma unroll
    for (int k = 3; k < 7; ++k) {
      const double dd = t[k + 1] - t[k];
      pq[k] = fma((13.0 / 3.0) * dd, dd, eps9);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) tc[k] = t[k + 4];
#pragma unroll
    for (int k = 0; k < 3; ++k) pc[k] = pq[k + 4];
    double T[7];  // cotangents of the intervals (j0 - 2 + k, j0 - 1 + k)
    T[0] = Tc0; T[1] = Tc1; T[2] = Tc2; T[3] = T[4] = T[5] = T[6] = 0.0;
    double o[4], gi[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      // state of the next cell, the face between, then the vector-Jacobian product of this cell
      Weno5State Sn = S;
      double uln = ul_right;
      if (m < 3 || more) {
        Sn = weno53_state(t[m + 1], t[m + 2], t[m + 3], t[m + 4], pq[m + 1], pq[m + 2], pq[m + 3]);
        uln = w[m + 3] + Sn.uL;
      }
      const double hG = hs * (v[m + 1] - v[m]);
      const LeanFace f = lean_face(hG, w[m + 2] + S.uR, uln, w[m + 2], w[m + 3]);
      o[m] = (fp.dp + f.dj) + (f.gR + fp.gL);
      weno53_vjp_acc(S, t[m], t[m + 1], t[m + 2], t[m + 3], f.gR, fp.gL, T[m], T[m + 1], T[m + 2], T[m + 3]);
      // the cell two behind has now received every cotangent of its two first differences
      const double om = (m == 0) ? o_m2 : ((m == 1) ? o_m1 : o[m - 2]);
      const double Tl = (m == 0) ? Tprev : T[m - 1];
      gi[m] = fma(1.0 / 6.0, Tl - T[m], om);
      S = Sn;
      fp = f;
    }
    // ---- results for the cells j0 - 2 .. j0 + 1 (reads of X are done: OUT may alias it)
    if (j0 > 0) {
      const double2 vm = V.ld2(j0 - 2);
      double r0 = fma(c_v, vm.x, gi[0]), r1 = fma(c_v, vm.y, gi[1]), r2 = fma(c_v, v[0], gi[2]), r3 = fma(c_v, v[1], gi[3]);
      if (MODE == 2) {
        const double2 a0 = A.ld2(j0 - 2), a1 = A.ld2(j0);
        r0 += a0.x; r1 += a0.y; r2 += a1.x; r3 += a1.y;
      }
      OUT.st2(j0 - 2, r0, r1);
      OUT.st2(j0, r2, r3);
      if (MODE == 1) {
        A.st2(j0 - 2, fma(0.75, r0, (1.0 / 3.0) * vm.x), fma(0.75, r1, (1.0 / 3.0) * vm.y));
        A.st2(j0, fma(0.75, r2, (1.0 / 3.0) * v[0]), fma(0.75, r3, (1.0 / 3.0) * v[1]));
      }
    } else {
      // the first two cells wait for the neighbour's cotangents: park what they need
      park[0 * 32] = T[0];  // (-2, -1): all of it goes to the left neighbour
      park[1 * 32] = T[1];  // (-1, 0)
      park[2 * 32] = T[2];  // (0, 1)
      park[3 * 32] = T[3];  // (1, 2): complete
      park[4 * 32] = o[0];
      park[5 * 32] = o[1];
    }
    Tprev = T[3];
    Tc0 = T[4]; Tc1 = T[5]; Tc2 = T[6];
    o_m2 = o[2];
    o_m1 = o[3];
  }
  // ---- the cotangents that belong to the neighbours' cells, and theirs that belong to mine
  const double Tm21 = park[0], Tm10 = park[32], T01 = park[64], T12 = park[96];
  const double fromR_a = __shfl_down_sync(kFull, Tm21, 1);  // right neighbour's (-2, -1) = my (C - 2, C - 1)
  const double fromR_b = __shfl_down_sync(kFull, Tm10, 1);  //                   (-1, 0)  = my (C - 1, C)
  const double fromL_a = __shfl_up_sync(kFull, Tc1, 1);     // left neighbour's (C - 1, C) = my (-1, 0)
  const double fromL_b = __shfl_up_sync(kFull, Tc2, 1);     //                  (C, C + 1) = my (0, 1)
  const double Ta = Tc0 + fromR_a, Tb = Tc1 + fromR_b;
  const double Tc = Tm10 + fromL_a, Td = T01 + fromL_b;
  const double g_m2 = fma(1.0 / 6.0, Tprev - Ta, o_m2);  // cell C - 2
  const double g_m1 = fma(1.0 / 6.0, Ta - Tb, o_m1);     // cell C - 1
  const double g_0 = fma(1.0 / 6.0, Tc - Td, park[4 * 32]);
  const double g_1 = fma(1.0 / 6.0, Td - T12, park[5 * 32]);
  const double2 v0 = V.ld2(0), vl = V.ld2(C - 2);
  double r0 = fma(c_v, v0.x, g_0), r1 = fma(c_v, v0.y, g_1), r2 = fma(c_v, vl.x, g_m2), r3 = fma(c_v, vl.y, g_m1);
  if (MODE == 2) {
    const double2 a0 = A.ld2(0), a1 = A.ld2(C - 2);
    r0 += a0.x; r1 += a0.y; r2 += a1.x; r3 += a1.y;
  }
  OUT.st2(0, r0, r1);
  OUT.st2(C - 2, r2, r3);
  if (MODE == 1) {
    A.st2(0, fma(0.75, r0, (1.0 / 3.0) * v0.x), fma(0.75, r1, (1.0 / 3.0) * v0.y));
    A.st2(C - 2, fma(0.75, r2, (1.0 / 3.0) * vl.x), fma(0.75, r3, (1.0 / 3.0) * vl.y));
  }
  OUT.st1(-1, __shfl_up_sync(kFull, r3, 1));
  OUT.st1(C, __shfl_down_sync(kFull, r0, 1));
}

// ---------------------------------------------------------------------------
// run cells -2 .. C + 1 of one global array into a window array (periodic images beyond the row ends)
// Dirichlet rows: the value of the window cell c < 0 or c >= n of a STATE array -- the boundary data of its stage in
// the three ghost cells, the outermost one repeated beyond them (cells no stored result depends on; finite filler)
__device__ __forceinline__ double rev_dirichlet_value(const double *__restrict__ gh, int c, int n) {
  int k = (c < 0) ? c + 3 : c - n + 3;
  k = k < 0 ? 0 : (k > 5 ? 5 : k);
  return gh[k];
}

// Dirichlet rows, lanes whose run reaches beyond the row: the cells beyond [0, n) of a state array take the data
// `gh` of its stage (apply_boundary at the top of the next right-hand side, schemes.py:343); of a cotangent array
// (gh == nullptr) they are zero -- the boundary data do not depend on the state
__device__ __forceinline__ void rev_dirichlet_fix(const int C, const LaneArray D, const double *__restrict__ gh, int r0,
                                                  int n) {
#pragma unroll 1
  for (int j = -2; j < C + 2; ++j) {
    const int c = r0 + j;
    if (c < 0 || c >= n) D.st1(j, gh != nullptr ? rev_dirichlet_value(gh, c, n) : 0.0);
  }
}

template <int CM>
__device__ __forceinline__ void rev_load_run(const int C, const double *__restrict__ src, int64_t base, int r0, int n,
                                             bool inside, const LaneArray D, int none_g = 0, bool dir = false,
                                             const double *__restrict__ gh = nullptr) {
  if (dir && !inside) {  // Dirichlet row end: interior cells as stored, boundary data (state) or zero (cotangent) beyond
#pragma unroll 1
    for (int j = -2; j < C + 2; ++j) {
      const int c = r0 + j;
      double v;
      if (c >= 0 && c < n) v = src[base + c];
      else v = (gh != nullptr) ? rev_dirichlet_value(gh, c, n) : 0.0;
      D.st1(j, v);
    }
    return;
  }
  if (none_g > 0 && !inside) {  // slab: stored ghost cells; further out never reaches a stored cell
#pragma unroll 1
    for (int j = -2; j < C + 2; ++j) {
      const int c = r0 + j;
      D.st1(j, (c >= -none_g && c < n + none_g) ? src[base + c] : 0.0);
    }
    return;
  }
  if (inside) {
    // four 16-byte loads in flight per iteration.  Measured (200 reverse steps of config 5): this form 242 ms; two
    // loads per iteration 248 ms (load latency exposed); fully unrolled and predicated over the longest run, every
    // load in flight, 254 ms (its code pushes the kernel past the 32 KB instruction cache)
#pragma unroll 4
    for (int k = 0; k < (C + 4) / 2; ++k) {
      const double2 a = *reinterpret_cast<const double2 *>(src + base + r0 - 2 + 2 * k);
      D.st2(-2 + 2 * k, a.x, a.y);
    }
  } else {
    int c = (r0 - 2) % n;
    if (c < 0) c += n;
#pragma unroll 1
    for (int j = -2; j < C + 2; ++j) {
      D.st1(j, src[base + c]);
      if (++c == n) c = 0;
    }
  }
}

template <int CM, int MINB, bool DIR = false>
__global__ void __launch_bounds__(32, MINB)
reverse_step_kernel(const RevParams p) {
  using Geo = RevGeometry<CM>;
  __shared__ double2 smem[Geo::kSmemDoubles / 2];
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int row = blockIdx.y + blockIdx.z * gridDim.y;
  const int n = p.n;
  const int C = (tile == p.tiles_per_row - 1) ? p.c_last : CM;  // cells per lane of this window
  const int window = 32 * C;
  const LaneArray P0{smem + lane}, P1{smem + Geo::kArrayDoubles / 2 + lane}, P2{smem + 2 * (Geo::kArrayDoubles / 2) + lane};
  double *park = reinterpret_cast<double *>(smem + 3 * (Geo::kArrayDoubles / 2)) + lane;
  const int64_t base = static_cast<int64_t>(row) * p.ld + p.g;
  const int r0 = tile * Geo::kEmit - kRevHalo + C * lane;  // ring coordinate of the lane's first cell
  const int none_g = p.bc_none ? p.g : 0;
  const bool inside = (r0 - 2 >= -none_g) && (r0 + C + 2 <= n + none_g);
  // Dirichlet rows: the boundary data of this row at the three stage times
  const double *gh0 = nullptr, *gh1 = nullptr, *gh2 = nullptr;
  if (DIR) {
    gh0 = p.ghost3 + static_cast<int64_t>(row) * p.ghost_ld;
    gh1 = gh0 + p.ghost_block;
    gh2 = gh1 + p.ghost_block;
  }
  const bool fix = DIR && !inside;

  rev_load_run<CM>(C, p.u, base, r0, n, inside, P0, none_g, DIR, gh0);
#ifndef PSK_HOST_EMU
  if (inside) {  // p' is needed after the recomputation: have its lines on their way
#pragma unroll 1
    for (int k = 0; k < C + 4; k += 16)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pin + base + r0 - 2 + k));
  }
#endif
  const double dt = p.dt[static_cast<int64_t>(row) * p.dt_stride];
  const double eps9 = p.eps * (1.0 / 9.0);
  const double cdt = (0.25 * p.invdx) * dt;  // the forward flux is scaled by 4 (psk_fast_kernels.cuh)

  // ---- recomputation of the stage values (timestepping.py:314-317): P0 = u, P1 = k1, P2 = k2
#pragma unroll 1
  for (int st = 0; st < 2; ++st) {
    rev_forward_stage(C, st == 1, st == 0 ? P0 : P1, st == 0 ? P1 : P2, P0, cdt, eps9);
    if (fix) rev_dirichlet_fix(C, st == 0 ? P1 : P2, st == 0 ? gh1 : gh2, r0, n);
  }
  if (p.dbg_k1 != nullptr) {
#pragma unroll 1
    for (int j = 0; j < C; ++j) {
      const int wi = C * lane + j, e = tile * Geo::kEmit + wi - kRevHalo;
      if (wi >= kRevHalo && wi < window - kRevHalo && e < n) {
        p.dbg_k1[base + e] = P1.ld1(j);
        p.dbg_k2[base + e] = P2.ld1(j);
      }
    }
  }
  // ---- the three adjoint stages (one body, executed three times)
  //   ph 0: P0 = p';  lam2 = 2/3 (p' + dt J(k2)^T p') over k2 (P2), P0 = 1/3 p' + 3/4 lam2
  //   ph 1:           lam1 = 1/4 (lam2 + dt J(k1)^T lam2) over k1 (P1)
  //   ph 2: P2 = u (an L2 hit);  p = P0 + lam1 + dt J(u)^T lam1 over P0
  const double hs = 0.5 * p.invdx * dt;
#pragma unroll 1
  for (int ph = 0; ph < 3; ++ph) {
    if (ph != 1)
      rev_load_run<CM>(C, ph == 0 ? p.pin : p.u, base, r0, n, inside, ph == 0 ? P0 : P2, none_g, DIR, ph == 0 ? nullptr : gh0);
    const double cv = (ph == 0) ? (2.0 / 3.0) : ((ph == 1) ? 0.25 : 1.0);
    const LaneArray X = (ph == 1) ? P1 : P2, V = (ph == 0) ? P0 : ((ph == 1) ? P2 : P1), OUT = (ph == 0) ? P2 : ((ph == 1) ? P1 : P0);
    rev_adjoint_stage(C, ph == 0 ? 1 : (ph == 1 ? 0 : 2), X, V, OUT, P0, cv, cv * hs, eps9, park);
    if (fix && ph != 2) rev_dirichlet_fix(C, OUT, nullptr, r0, n);  // no cotangent lives on the boundary data
  }

  // ---- p of the stored window cells
  const int e0 = tile * Geo::kEmit + C * lane - kRevHalo;  // interior coordinate of the lane's first cell if stored
#pragma unroll 1
  for (int j = 0; j < C; j += 2) {
    const int wi = C * lane + j;
    if (wi >= kRevHalo && wi < window - kRevHalo && e0 + j < n) {
      const double2 q = P0.ld2(j);
      if (e0 + j + 1 < n) {
        *reinterpret_cast<double2 *>(p.pout + base + e0 + j) = q;
      } else {
        p.pout[base + e0 + j] = q.x;
      }
    }
  }
}

}  // namespace psk
