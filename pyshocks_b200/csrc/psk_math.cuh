// psk_math.cuh -- per-cell device arithmetic of the pyshocks hot path (fp64, sm_100a).
//
// Two arithmetic contracts (include/psk.h, enum psk_math):
//   STRICT  the reference's operation order (as NumPy evaluates it; see
//           oracle/psk_oracle.c), written with the non-contracting intrinsics
//           __dadd_rn / __dmul_rn / __ddiv_rn so that nvcc forms no FMA:
//           bit-identical to the C oracle.
//   FAST    same algorithm re-associated for the FP64 pipe: first/second
//           differences shared between neighbouring cells, smoothness
//           indicators shared between the left and right value of a cell,
//           omega_k ~ d_k * prod_{l != k} (eps + beta_l)^2 so that each
//           reconstructed value costs one reciprocal instead of six divisions.
//
// Reference citations are relative to /root/reference/src/pyshocks.
#pragma once

#include <cstdint>

#include "../../include/psk.h"

namespace psk {

// ---------------------------------------------------------------------------
// strict helpers: never contracted, IEEE round-to-nearest

__device__ __forceinline__ double sadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ssub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double smul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double sdiv(double a, double b) { return __ddiv_rn(a, b); }

// jnp.maximum / jnp.minimum propagate NaN
__device__ __forceinline__ double max_nan(double x, double y) { return (x > y || x != x) ? x : y; }
__device__ __forceinline__ double min_nan(double x, double y) { return (x < y || x != x) ? x : y; }

// ---------------------------------------------------------------------------
// WENO-JS one-sided values, STRICT (weno.py:114-157, :166-256; reconstruction.py:351-355)
// (m2, m1, c, p1, p2) = u[i-2 .. i+2]; returns the value at the right face of cell i.
// The left-face value is the same function of the reversed arguments
// (reconstruction.py:374-375).

__device__ __forceinline__ double weno53_side_strict(double m2, double m1, double c, double p1,
                                                      double p2, double eps) {
  const double a0 = 13.0 / 12.0, a1 = 1.0 / 4.0;
  // np.convolve accumulates over ascending memory index (u[i-2] first)
  double c00 = sadd(sadd(smul(m2, 1.0), smul(m1, -2.0)), smul(c, 1.0));
  double c01 = sadd(sadd(smul(m2, 1.0), smul(m1, -4.0)), smul(c, 3.0));
  double c10 = sadd(sadd(smul(m1, 1.0), smul(c, -2.0)), smul(p1, 1.0));
  double c11 = sadd(sadd(smul(m1, -1.0), smul(c, 0.0)), smul(p1, 1.0));
  double c20 = sadd(sadd(smul(c, 1.0), smul(p1, -2.0)), smul(p2, 1.0));
  double c21 = sadd(sadd(smul(c, 3.0), smul(p1, -4.0)), smul(p2, 1.0));
  double b0 = sadd(smul(a0, smul(c00, c00)), smul(a1, smul(c01, c01)));
  double b1 = sadd(smul(a0, smul(c10, c10)), smul(a1, smul(c11, c11)));
  double b2 = sadd(smul(a0, smul(c20, c20)), smul(a1, smul(c21, c21)));
  double q0 = sadd(sadd(smul(m2, 2.0 / 6.0), smul(m1, -7.0 / 6.0)), smul(c, 11.0 / 6.0));
  double q1 = sadd(sadd(smul(m1, -1.0 / 6.0), smul(c, 5.0 / 6.0)), smul(p1, 2.0 / 6.0));
  double q2 = sadd(sadd(smul(c, 2.0 / 6.0), smul(p1, 5.0 / 6.0)), smul(p2, -1.0 / 6.0));
  double e0 = sadd(eps, b0), e1 = sadd(eps, b1), e2 = sadd(eps, b2);
  double al0 = sdiv(1.0 / 10.0, smul(e0, e0));
  double al1 = sdiv(6.0 / 10.0, smul(e1, e1));
  double al2 = sdiv(3.0 / 10.0, smul(e2, e2));
  double tot = sadd(sadd(al0, al1), al2);
  return sadd(sadd(smul(sdiv(al0, tot), q0), smul(sdiv(al1, tot), q1)), smul(sdiv(al2, tot), q2));
}

__device__ __forceinline__ double weno32_side_strict(double m1, double c, double p1, double eps) {
  double c0 = sadd(smul(m1, -1.0), smul(c, 1.0));
  double c1 = sadd(sadd(smul(m1, 0.0), smul(c, -1.0)), smul(p1, 1.0));
  double b0 = smul(1.0, smul(c0, c0));
  double b1 = smul(1.0, smul(c1, c1));
  double q0 = sadd(smul(m1, -1.0 / 2.0), smul(c, 3.0 / 2.0));
  double q1 = sadd(sadd(smul(m1, 0.0), smul(c, 1.0 / 2.0)), smul(p1, 1.0 / 2.0));
  double e0 = sadd(eps, b0), e1 = sadd(eps, b1);
  double al0 = sdiv(1.0 / 3.0, smul(e0, e0));
  double al1 = sdiv(2.0 / 3.0, smul(e1, e1));
  double tot = sadd(al0, al1);
  return sadd(smul(sdiv(al0, tot), q0), smul(sdiv(al1, tot), q1));
}

// ---------------------------------------------------------------------------
// FAST reciprocal: MUFU.RCP64H seed (rel. error <= 2^-23) refined by one cubic
// step, y = y0 + y0 (e + e^2), e = 1 - x y0  ->  rel. error ~ 2^-69 before the
// final rounding, i.e. a result within ~1 ulp.  4 FP64-pipe instructions instead
// of the ~10 + slow-path branch of an IEEE division.  Valid for normal, finite,
// non-zero x (the WENO denominators are >= eps^4 > 0).
__device__ __forceinline__ double fast_rcp(double x) {
  double y0;
#ifdef PSK_HOST_EMU  // host build of the kernels for the CPU tests (tests/host/emu/cuda_runtime.h)
  y0 = psk_emu_rcp_seed(x);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
#endif
  double e = fma(-x, y0, 1.0);
  double t = fma(e, e, e);
  return fma(y0, t, y0);
}

// FAST WENO-JS5 pair at one cell from pre-differenced data.
//   c                    cell value u[i]
//   dm2, dm1, dp0, dp1   HALF first differences 0.5 (u[i-1]-u[i-2]), 0.5 (u[i]-u[i-1]),
//                        0.5 (u[i+1]-u[i]), 0.5 (u[i+2]-u[i+1])
//   pm1, p0, pp1         13/12 (second difference)^2 centred at i-1, i, i+1
// Smoothness indicators (weno.py:218-231):
//   beta0 = 13/12 (c-2b+a)^2 + 1/4 (3c-4b+a)^2,  3c-4b+a = 3 (c-b) - (b-a)
//   beta1 = 13/12 (d-2c+b)^2 + 1/4 (d-b)^2,      d-b     = (d-c) + (c-b)
//   beta2 = 13/12 (e-2d+c)^2 + 1/4 (e-4d+3c)^2,  e-4d+3c = (e-d) - 3 (d-c)
// are the same three numbers for the right and the left value of the cell
// (mirror symmetry), so they are computed once.
struct Weno5Pair {
  double ul, ur;
};

__device__ __forceinline__ Weno5Pair weno53_pair_fast(double c, double dm2, double dm1, double dp0,
                                                       double dp1, double pm1, double p0,
                                                       double pp1, double eps) {
  double s0 = fma(3.0, dm1, -dm2);
  double s1 = dp0 + dm1;
  double s2 = fma(-3.0, dp0, dp1);
  double e0 = fma(s0, s0, pm1) + eps;
  double e1 = fma(s1, s1, p0) + eps;
  double e2 = fma(s2, s2, pp1) + eps;
  // alpha_k = d_k / e_k^2  ==  d_k (e_l e_m)^2 / (e_0 e_1 e_2)^2 ; the common factor cancels
  double e12 = e1 * e2, e02 = e0 * e2, e01 = e0 * e1;
  double w0 = e12 * e12, w1 = e02 * e02, w2 = e01 * e01;  // weights of stencils 0, 1, 2 (un-normalised, without d_k)
  double a1 = 0.6 * w1;
  double aR0 = 0.1 * w0, aR2 = 0.3 * w2;  // right value: d = (0.1, 0.6, 0.3), weno.py:244
  double aL0 = 0.3 * w0, aL2 = 0.1 * w2;  // left value: mirrored ideal weights
  // candidate values minus c (weno.py:234-241), in half differences:
  //   right: q0 - c = ( 5 (c-b) - 2 (b-a)) / 6,  q1 - c = (2 (d-c) + (c-b)) / 6,  q2 - c = (4 (d-c) - (e-d)) / 6
  //   left : mirror images
  const double k13 = 1.0 / 3.0, k23 = 2.0 / 3.0, k43 = 4.0 / 3.0, k53 = 5.0 / 3.0;
  double rR0 = fma(k53, dm1, -k23 * dm2);
  double rR1 = fma(k23, dp0, k13 * dm1);
  double rR2 = fma(k43, dp0, -k13 * dp1);
  double rL2 = fma(-k53, dp0, k23 * dp1);   // stencil {i, i+1, i+2} -> left face
  double rL1 = fma(-k23, dm1, -k13 * dp0);  // stencil {i-1, i, i+1}
  double rL0 = fma(-k43, dm1, k13 * dm2);   // stencil {i-2, i-1, i}
  double numR = fma(aR2, rR2, fma(a1, rR1, aR0 * rR0));
  double numL = fma(aL2, rL2, fma(a1, rL1, aL0 * rL0));
  double denR = (aR0 + a1) + aR2;
  double denL = (aL0 + a1) + aL2;
  Weno5Pair out;
  out.ur = fma(numR, fast_rcp(denR), c);
  out.ul = fma(numL, fast_rcp(denL), c);
  return out;
}

// The pair used by every FAST WENO-JS5 stage kernel.  All quantities are in units of SIXTHS of the
// first differences, t = (u[k+1] - u[k]) / 6, which turns the candidate offsets into small-integer
// combinations,
//   right: q0 - c = 5 t(-1) - 2 t(-2), q1 - c = 2 t(0) + t(-1), q2 - c = 4 t(0) - t(+1)
//   left : q0 - c = t(-2) - 4 t(-1),   q1 - c = -(2 t(-1) + t(0)), q2 - c = 2 t(+1) - 5 t(0)
// and the smoothness indicators into beta_k / 9 = (13/3) (second difference of t)^2 + s_k^2,
// s0 = 3 t(-1) - t(-2), s1 = t(0) + t(-1), s2 = t(+1) - 3 t(0).  The common factor 9 and the
// common factor 1/10 of the ideal weights (1, 6, 3)/10 cancel in the weights, so eps is
// passed pre-divided (pm1, p0, pp1 = (13/3) dd^2 + eps / 9).  The candidate offsets are written
// in the smoothness stencils' own linear forms (one instruction each):
//   right: q0 - c = 2 s0 - t(-1),  q1 - c = 2 t(0) + t(-1),   q2 - c = t(0) - s2
//   left : q0 - c = -(s0 + t(-1)), q1 - c = -(2 t(-1) + t(0)), q2 - c = 2 s2 + t(0)
// and the ideal-weight factors 3 are folded into the offsets instead of the weights (37 FP64
// instructions for the two values of a cell):
//   right: N = W0 r0 + (6 W1) r1 + W2 (3 r2),   D = W0 + 6 W1 + 3 W2,   3 r2 = 3 (t(0) - s2) = t(+1) - 4 s2
//   left : N = W0 (3 r0) + (6 W1) r1 + W2 r2,   D = 3 W0 + 6 W1 + W2,   3 r0 = -3 (s0 + t(-1)) = -(4 s0 + t(-2))
// (3 t(0) = t(+1) - s2 and 3 t(-1) = s0 + t(-2) by the definitions of s2 and s0).  The one product
// that feeds plain additions is written with __dmul_rn so that no compiler decision about
// contraction can make the arithmetic of a cell depend on where the cell sits in a warp.
__device__ __forceinline__ Weno5Pair weno53_pair_lean(double c, double tm2, double tm1, double tp0,
                                                       double tp1, double pm1, double p0,
                                                       double pp1) {
  const double s0 = fma(3.0, tm1, -tm2);
  const double s1 = tp0 + tm1;
  const double s2 = fma(-3.0, tp0, tp1);
  const double e0 = fma(s0, s0, pm1);
  const double e1 = fma(s1, s1, p0);
  const double e2 = fma(s2, s2, pp1);
  const double e12 = e1 * e2, e02 = e0 * e2, e01 = e0 * e1;
  const double w0 = e12 * e12, w2 = e01 * e01;
  const double a1 = __dmul_rn(6.0, e02 * e02);
  const double rR0 = fma(2.0, s0, -tm1);
  const double rR1 = fma(2.0, tp0, tm1);
  const double rR2x3 = fma(-4.0, s2, tp1);
  const double nL0x3 = fma(4.0, s0, tm2);  // -(3 r0), left
  const double nL1 = fma(2.0, tm1, tp0);   // -(r1), left
  const double rL2 = fma(2.0, s2, tp0);
  const double numR = fma(w2, rR2x3, fma(a1, rR1, w0 * rR0));
  const double numL = fma(w2, rL2, -fma(a1, nL1, w0 * nL0x3));
  const double denR = fma(3.0, w2, w0 + a1);
  const double denL = fma(3.0, w0, a1 + w2);
  Weno5Pair out;
  out.ur = fma(numR, fast_rcp(denR), c);
  out.ul = fma(numL, fast_rcp(denL), c);
  return out;
}

// FAST pair straight from the five cell values (used where no sliding window exists)
__device__ __forceinline__ Weno5Pair weno53_pair_fast_cells(double m2, double m1, double c, double p1,
                                                             double p2, double eps) {
  double dm2 = 0.5 * (m1 - m2), dm1 = 0.5 * (c - m1), dp0 = 0.5 * (p1 - c), dp1 = 0.5 * (p2 - p1);
  const double k = 4.0 * (13.0 / 12.0);  // second difference of half differences is half the true one
  double tm1 = dm1 - dm2, t0 = dp0 - dm1, tp1 = dp1 - dp0;
  return weno53_pair_fast(c, dm2, dm1, dp0, dp1, k * tm1 * tm1, k * t0 * t0, k * tp1 * tp1, eps);
}

__device__ __forceinline__ double weno32_side_fast(double m1, double c, double p1, double eps) {
  double d0 = c - m1, d1 = p1 - c;
  double e0 = fma(d0, d0, eps), e1 = fma(d1, d1, eps);
  double w0 = e1 * e1, w1 = 2.0 * (e0 * e0);  // alpha_0 : alpha_1 = (1/3)/e0^2 : (2/3)/e1^2
  // q0 - c = (c - m1)/2, q1 - c = (p1 - c)/2
  double num = fma(w1, d1, w0 * d0);
  return fma(0.5 * num, fast_rcp(w0 + w1), c);
}

// ---------------------------------------------------------------------------
// ESWENO32 (reconstruction.py:386-439, weno.py:284-296): the JS-3 stencils with
//   alpha_k = d_k (1 + tau / (eps + beta_k)),  tau = (u[i+1] - 2 u[i] + u[i-1])^2,
// tau being ZERO at the two ends of the array (jnp.pad of the interior expression), which only
// the ghost-row / parity kernels ever see (tau_zero).  om0 is omega_0 of the right-value
// weights, the quantity the dissipative flux of the Burgers ESWENO32 scheme is built from
// (burgers/schemes.py:237-247).
struct EsCell {
  double ul, ur, om0;
};

__device__ __forceinline__ double esweno32_side_strict(double m1, double c, double p1, double eps,
                                                       bool tau_zero, double *om0) {
  double c0 = sadd(smul(m1, -1.0), smul(c, 1.0));
  double c1 = sadd(sadd(smul(m1, 0.0), smul(c, -1.0)), smul(p1, 1.0));
  double b0 = smul(1.0, smul(c0, c0));
  double b1 = smul(1.0, smul(c1, c1));
  double q0 = sadd(smul(m1, -1.0 / 2.0), smul(c, 3.0 / 2.0));
  double q1 = sadd(sadd(smul(m1, 0.0), smul(c, 1.0 / 2.0)), smul(p1, 1.0 / 2.0));
  double tt = sadd(ssub(p1, smul(2.0, c)), m1);
  double tau = tau_zero ? 0.0 : smul(tt, tt);
  double al0 = smul(1.0 / 3.0, sadd(1.0, sdiv(tau, sadd(eps, b0))));
  double al1 = smul(2.0 / 3.0, sadd(1.0, sdiv(tau, sadd(eps, b1))));
  double tot = sadd(al0, al1);
  if (om0 != nullptr) *om0 = sdiv(al0, tot);
  return sadd(smul(sdiv(al0, tot), q0), smul(sdiv(al1, tot), q1));
}

template <bool STRICT>
__device__ __forceinline__ EsCell esweno32_cell(double m1, double c, double p1, double eps,
                                                bool tau_zero) {
  EsCell o;
  if (STRICT) {
    o.ur = esweno32_side_strict(m1, c, p1, eps, tau_zero, &o.om0);
    o.ul = esweno32_side_strict(p1, c, m1, eps, tau_zero, nullptr);
    return o;
  }
  // FAST: with e_k = eps + beta_k, A0 = (e0 + tau) e1, A1 = (e1 + tau) e0:
  //   right  omega = (A0, 2 A1) / (A0 + 2 A1),  q - c = (d0, d1) / 2
  //   left   omega = (A1, 2 A0) / (A1 + 2 A0),  q - c = (-d1, -d0) / 2
  const double d0 = c - m1, d1 = p1 - c;
  const double e0 = fma(d0, d0, eps), e1 = fma(d1, d1, eps);
  const double tt = d1 - d0;
  const double tau = tau_zero ? 0.0 : tt * tt;
  const double A0 = (e0 + tau) * e1, A1 = (e1 + tau) * e0;
  const double iR = fast_rcp(fma(2.0, A1, A0)), iL = fast_rcp(fma(2.0, A0, A1));
  o.om0 = A0 * iR;
  o.ur = fma(0.5 * fma(A0, d0, 2.0 * (A1 * d1)), iR, c);
  o.ul = fma(-0.5 * fma(A1, d1, 2.0 * (A0 * d0)), iL, c);
  return o;
}

// dissipative flux of the Burgers ESWENO32 scheme at the face between cells j and j+1
// (burgers/schemes.py:243-247): mu = sqrt((om_p - om_j)^2 + delta^2) / 8,
// g = -(mu + (om_p - om_j) / 8) (w_p - w_j).  FAST returns 2 g (the scale of the upwind flux).
template <bool STRICT>
__device__ __forceinline__ double esweno_gnum(double omj, double omp, double wj, double wp, double delta) {
  if (STRICT) {
    const double dom = ssub(omp, omj);
    const double mu = sdiv(__dsqrt_rn(sadd(smul(dom, dom), smul(delta, delta))), 8.0);
    return smul(-sadd(mu, sdiv(dom, 8.0)), ssub(wp, wj));
  }
  const double dom = omp - omj;
  return -0.25 * (sqrt(fma(dom, dom, delta * delta)) + dom) * (wp - wj);
}

// ---------------------------------------------------------------------------
// generic per-cell reconstruction (no sliding window): (ul, ur) of cell with
// stencil values v[-2..2] (v points at the cell); used by the edge/naive kernels
// and by the tile kernel for STRICT math and the non-JS5 reconstructions.

template <int REC, bool STRICT>
__device__ __forceinline__ Weno5Pair reconstruct_cell(double m2, double m1, double c, double p1,
                                                       double p2, double eps, bool tau_zero = false) {
  Weno5Pair o;
  if (REC == PSK_REC_CONSTANT) {
    o.ul = c;
    o.ur = c;  // reconstruction.py:153-163
  } else if (REC == PSK_REC_ESWENO32) {
    const EsCell e = esweno32_cell<STRICT>(m1, c, p1, eps, tau_zero);
    o.ul = e.ul;
    o.ur = e.ur;
  } else if (REC == PSK_REC_WENOJS32) {
    if (STRICT) {
      o.ur = weno32_side_strict(m1, c, p1, eps);
      o.ul = weno32_side_strict(p1, c, m1, eps);
    } else {
      o.ur = weno32_side_fast(m1, c, p1, eps);
      o.ul = weno32_side_fast(p1, c, m1, eps);
    }
  } else {
    if (STRICT) {
      o.ur = weno53_side_strict(m2, m1, c, p1, p2, eps);
      o.ul = weno53_side_strict(p2, p1, c, m1, m2, eps);
    } else {
      o = weno53_pair_fast_cells(m2, m1, c, p1, p2, eps);
    }
  }
  return o;
}

// ---------------------------------------------------------------------------
// numerical flux at the face between cells j (left) and j+1 (right)
//   urj  right-face value of cell j        ulp  left-face value of cell j+1
//   wj, wp  cell values (after the boundary condition)
//   speed   global max |w| for Lax-Friedrichs (scalar.py:277), nu: scalar.py:231-234
//   arj, alp  reconstructed velocity for advection / continuity

template <int EQ, int FLUX, bool STRICT>
__device__ __forceinline__ double face_flux(double urj, double ulp, double wj, double wp,
                                             double speed, double nu, double arj, double alp) {
  if (EQ == PSK_EQ_BURGERS) {
    if (FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
      // scalar.py:231-249: 0.5 (f(ul) + f(ur)) - 0.5 a nu (ul - ur), f = u^2 / 2
      double a = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? speed : max_nan(fabs(wp), fabs(wj));
      if (STRICT) {
        double fr = sdiv(smul(urj, urj), 2.0), fl = sdiv(smul(ulp, ulp), 2.0);
        return ssub(smul(0.5, sadd(fl, fr)), smul(smul(smul(0.5, a), nu), ssub(ulp, urj)));
      }
      double ss = fma(urj, urj, ulp * ulp);
      return fma(-0.5 * (a * nu), ulp - urj, 0.25 * ss);
    }
    if (FLUX == PSK_FLUX_UPWIND || FLUX == PSK_FLUX_ESWENO) {
      // scalar.py:123-132 with a = u: where((ur[j] + ul[j+1]) / 2 > 0, f(ur[j]), f(ul[j+1]));
      // the ESWENO32 scheme adds esweno_gnum to this (burgers/schemes.py:255-256)
      double aavg = STRICT ? sdiv(sadd(urj, ulp), 2.0) : 0.5 * (urj + ulp);
      double v = aavg > 0.0 ? urj : ulp;
      return STRICT ? sdiv(smul(v, v), 2.0) : 0.5 * (v * v);
    }
    // Engquist-Osher, omega = 0 (scalar.py:311-322, burgers/schemes.py:184)
    double vp = max_nan(urj, 0.0), vm = min_nan(ulp, 0.0);
    if (STRICT)
      return ssub(sadd(sdiv(smul(vp, vp), 2.0), sdiv(smul(vm, vm), 2.0)), 0.0);
    return 0.5 * fma(vp, vp, vm * vm);
  }
  // advection/schemes.py:107-114, continuity/schemes.py:103-110
  double aavg = STRICT ? sdiv(sadd(arj, alp), 2.0) : 0.5 * (arj + alp);
  if (EQ == PSK_EQ_ADVECTION) return aavg > 0.0 ? urj : ulp;
  return aavg > 0.0 ? (STRICT ? smul(arj, urj) : arj * urj) : (STRICT ? smul(alp, ulp) : alp * ulp);
}

// RHS of cell i from its two face fluxes (schemes.py:346, advection/schemes.py:73)
template <int EQ, bool STRICT>
__device__ __forceinline__ double rhs_from_faces(double f_lo, double f_hi, double vel, double dx,
                                                  double invdx) {
  if (STRICT) {
    double df = ssub(f_hi, f_lo);
    if (EQ == PSK_EQ_ADVECTION) return sdiv(smul(-vel, df), dx);
    return sdiv(-df, dx);
  }
  double df = f_lo - f_hi;
  if (EQ == PSK_EQ_ADVECTION) return (vel * df) * invdx;
  return df * invdx;
}

// SSPRK33 stage combine (timestepping.py:313-320)
//   stage 0: L ; 1: w + dt L ; 2: 3/4 u0 + 1/4 (w + dt L) ; 3: 1/3 u0 + 2/3 (w + dt L)
template <bool STRICT>
__device__ __forceinline__ double stage_combine(int stage, double u0, double w, double dt, double L) {
  if (stage == 0) return L;
  if (STRICT) {
    double k = sadd(w, smul(dt, L));
    if (stage == 1) return k;
    if (stage == 2) return sadd(smul(3.0 / 4.0, u0), smul(1.0 / 4.0, k));
    return sadd(smul(1.0 / 3.0, u0), smul(2.0 / 3.0, k));
  }
  double k = fma(dt, L, w);
  if (stage == 1) return k;
  if (stage == 2) return fma(0.25, k, 0.75 * u0);
  return fma(2.0 / 3.0, k, (1.0 / 3.0) * u0);
}

}  // namespace psk

// ===========================================================================
// Discrete adjoint (vector-Jacobian products) of the pieces above.
//
// The reference differentiates `advance` with jax.jacfwd (timestepping.py:174); these
// are the hand-derived transposes of the same operations, with JAX's conventions at
// kinks: where() differentiates the selected branch only, maximum/minimum split ties
// 1/2-1/2, abs'(0) = 0.
namespace psk {

// Derivative of a WENO-JS value U = sum_k omega_k q_k,
//   omega_k = alpha_k / sum(alpha), alpha_k = d_k / e_k^2, e_k = eps + beta_k:
//   dU = sum_k omega_k dq_k + sum_k gamma_k dbeta_k,  gamma_k = -2 omega_k (q_k - U) / e_k.
// (q_k - U) is formed from the deviations r_k = q_k - c, so no cancellation against the
// cell value enters.

struct Weno5Vjp {
  double d[5];  // contributions to the cotangents of u[i-2 .. i+2]
};

// cotangents g_ur, g_ul of the right / left face value of the cell (m2, m1, c, p1, p2)
__device__ __forceinline__ Weno5Vjp weno53_pair_vjp(double m2, double m1, double c, double p1,
                                                    double p2, double eps, double g_ur,
                                                    double g_ul) {
  // smoothness indicators (weno.py:218-231)
  const double t0 = (m2 - 2.0 * m1) + c, s0 = (m2 - 4.0 * m1) + 3.0 * c;
  const double t1 = (m1 - 2.0 * c) + p1, s1 = p1 - m1;
  const double t2 = (c - 2.0 * p1) + p2, s2 = (3.0 * c - 4.0 * p1) + p2;
  const double k13 = 13.0 / 12.0;
  const double e0 = fma(k13 * t0, t0, fma(0.25 * s0, s0, eps));
  const double e1 = fma(k13 * t1, t1, fma(0.25 * s1, s1, eps));
  const double e2 = fma(k13 * t2, t2, fma(0.25 * s2, s2, eps));
  const double i0 = fast_rcp(e0), i1 = fast_rcp(e1), i2 = fast_rcp(e2);
  const double w0 = i0 * i0, w1 = i1 * i1, w2 = i2 * i2;  // 1 / e_k^2
  // right value: d = (0.1, 0.6, 0.3) on stencils (0, 1, 2); left value: (0.3, 0.6, 0.1)
  const double aR0 = 0.1 * w0, a1 = 0.6 * w1, aR2 = 0.3 * w2;
  const double aL0 = 0.3 * w0, aL2 = 0.1 * w2;
  const double iR = fast_rcp((aR0 + a1) + aR2), iL = fast_rcp((aL0 + a1) + aL2);
  const double oR0 = aR0 * iR, oR1 = a1 * iR, oR2 = aR2 * iR;
  const double oL0 = aL0 * iL, oL1 = a1 * iL, oL2 = aL2 * iL;
  // candidate values minus c (weno.py:234-241 and their mirror images)
  const double dm2 = m1 - m2, dm1 = c - m1, dp0 = p1 - c, dp1 = p2 - p1;
  const double k6 = 1.0 / 6.0;
  const double rR0 = k6 * (5.0 * dm1 - 2.0 * dm2), rR1 = k6 * (2.0 * dp0 + dm1),
               rR2 = k6 * (4.0 * dp0 - dp1);
  const double rL0 = k6 * (dm2 - 4.0 * dm1), rL1 = -k6 * (2.0 * dm1 + dp0),
               rL2 = k6 * (2.0 * dp1 - 5.0 * dp0);
  const double uR = fma(oR2, rR2, fma(oR1, rR1, oR0 * rR0));  // U_R - c
  const double uL = fma(oL2, rL2, fma(oL1, rL1, oL0 * rL0));  // U_L - c
  // cotangents on beta_k, both sides folded together
  const double G0 = -2.0 * i0 * (g_ur * oR0 * (rR0 - uR) + g_ul * oL0 * (rL0 - uL));
  const double G1 = -2.0 * i1 * (g_ur * oR1 * (rR1 - uR) + g_ul * oL1 * (rL1 - uL));
  const double G2 = -2.0 * i2 * (g_ur * oR2 * (rR2 - uR) + g_ul * oL2 * (rL2 - uL));
  // d beta_k = (13/6) t_k dt_k + (1/2) s_k ds_k
  const double k136 = 13.0 / 6.0;
  const double A0 = G0 * k136 * t0, B0 = G0 * 0.5 * s0;
  const double A1 = G1 * k136 * t1, B1 = G1 * 0.5 * s1;
  const double A2 = G2 * k136 * t2, B2 = G2 * 0.5 * s2;
  // cotangents on the candidate values
  const double hR0 = g_ur * oR0 * k6, hR1 = g_ur * oR1 * k6, hR2 = g_ur * oR2 * k6;
  const double hL0 = g_ul * oL0 * k6, hL1 = g_ul * oL1 * k6, hL2 = g_ul * oL2 * k6;
  Weno5Vjp o;
  // stencil taps: t0,s0 on (m2,m1,c) = (1,-2,1),(1,-4,3); t1,s1 on (m1,c,p1) = (1,-2,1),(-1,0,1);
  // t2,s2 on (c,p1,p2) = (1,-2,1),(3,-4,1); qR0 (2,-7,11), qR1 (-1,5,2), qR2 (2,5,-1) [/6];
  // qL0 on (m2,m1,c) = (-1,5,2), qL1 on (m1,c,p1) = (2,5,-1), qL2 on (c,p1,p2) = (11,-7,2) [/6]
  o.d[0] = (A0 + B0) + fma(2.0, hR0, -hL0);
  o.d[1] = (fma(-2.0, A0, -4.0 * B0) + (A1 - B1)) + (fma(-7.0, hR0, -hR1) + fma(5.0, hL0, 2.0 * hL1));
  o.d[2] = ((fma(3.0, B0, A0) - 2.0 * A1) + fma(3.0, B2, A2)) +
           ((fma(11.0, hR0, 5.0 * hR1) + 2.0 * hR2) + (fma(2.0, hL0, 5.0 * hL1) + 11.0 * hL2));
  o.d[3] = ((A1 + B1) + fma(-2.0, A2, -4.0 * B2)) + (fma(2.0, hR1, 5.0 * hR2) + fma(-7.0, hL2, -hL1));
  o.d[4] = (A2 + B2) + fma(2.0, hL2, -hR2);
  return o;
}

// JS-3 (weno.py:166-203): stencils {i-1, i} and {i, i+1}
__device__ __forceinline__ Weno5Vjp weno32_pair_vjp(double m1, double c, double p1, double eps,
                                                    double g_ur, double g_ul) {
  const double d0 = c - m1, d1 = p1 - c;
  const double e0 = fma(d0, d0, eps), e1 = fma(d1, d1, eps);
  const double i0 = fast_rcp(e0), i1 = fast_rcp(e1);
  const double w0 = i0 * i0, w1 = i1 * i1;
  // right: alpha = (1/3 w0, 2/3 w1), q0 - c = d0 / 2, q1 - c = d1 / 2
  // left (mirror): alpha = (2/3 w0, 1/3 w1), q0 - c = -d0 / 2 [stencil {i-1,i}], q1 - c = -d1 / 2
  const double aR0 = (1.0 / 3.0) * w0, aR1 = (2.0 / 3.0) * w1;
  const double aL0 = (2.0 / 3.0) * w0, aL1 = (1.0 / 3.0) * w1;
  const double iR = fast_rcp(aR0 + aR1), iL = fast_rcp(aL0 + aL1);
  const double oR0 = aR0 * iR, oR1 = aR1 * iR, oL0 = aL0 * iL, oL1 = aL1 * iL;
  const double rR0 = 0.5 * d0, rR1 = 0.5 * d1;
  // left value at x_{i-1/2}: stencil {i-1,i}: (m1 + c)/2 -> -d0/2 ; stencil {i,i+1}: (3c - p1)/2 -> -d1/2
  const double rL0 = -0.5 * d0, rL1 = -0.5 * d1;
  const double uR = fma(oR1, rR1, oR0 * rR0), uL = fma(oL1, rL1, oL0 * rL0);
  const double G0 = -2.0 * i0 * (g_ur * oR0 * (rR0 - uR) + g_ul * oL0 * (rL0 - uL));
  const double G1 = -2.0 * i1 * (g_ur * oR1 * (rR1 - uR) + g_ul * oL1 * (rL1 - uL));
  // d beta_0 = 2 d0 (dc - dm1), d beta_1 = 2 d1 (dp1 - dc)
  const double A0 = 2.0 * G0 * d0, A1 = 2.0 * G1 * d1;
  // candidates: qR0 = -m1/2 + 3c/2, qR1 = c/2 + p1/2, qL0 = m1/2 + c/2, qL1 = 3c/2 - p1/2
  const double hR0 = 0.5 * g_ur * oR0, hR1 = 0.5 * g_ur * oR1;
  const double hL0 = 0.5 * g_ul * oL0, hL1 = 0.5 * g_ul * oL1;
  Weno5Vjp o;
  o.d[0] = 0.0;
  o.d[1] = -A0 + (hL0 - hR0);
  o.d[2] = (A0 - A1) + ((3.0 * hR0 + hR1) + (hL0 + 3.0 * hL1));
  o.d[3] = A1 + (hR1 - hL1);
  o.d[4] = 0.0;
  return o;
}

// The same vector-Jacobian product in the sixths formulation of the warp kernels: inputs are
// t(-2..1) = first differences / 6 around the cell and the three (13/3) dd^2 + eps/9 terms;
// everything is linear in the cotangents of the four t's, which are then spread onto the
// five cells.  One reciprocal of e0 e1 e2 gives the three 1/e_k.
__device__ __forceinline__ Weno5Vjp weno53_pair_vjp_sixths(double tm2, double tm1, double tp0,
                                                           double tp1, double pm1, double p0,
                                                           double pp1, double g_ur, double g_ul) {
  const double s0 = fma(3.0, tm1, -tm2);
  const double s1 = tp0 + tm1;
  const double s2 = fma(-3.0, tp0, tp1);
  const double e0 = fma(s0, s0, pm1);
  const double e1 = fma(s1, s1, p0);
  const double e2 = fma(s2, s2, pp1);
  const double e12 = e1 * e2, e02 = e0 * e2, e01 = e0 * e1;
  const double w0 = e12 * e12, w1 = e02 * e02, w2 = e01 * e01;
  const double a1 = 6.0 * w1, a2R = 3.0 * w2, a0L = 3.0 * w0;
  const double iR = fast_rcp((w0 + a1) + a2R), iL = fast_rcp((a0L + a1) + w2);
  const double oR0 = w0 * iR, oR1 = a1 * iR, oR2 = a2R * iR;
  const double oL0 = a0L * iL, oL1 = a1 * iL, oL2 = w2 * iL;
  const double rR0 = fma(2.0, s0, -tm1), rR1 = fma(2.0, tp0, tm1), rR2 = tp0 - s2;
  const double rL0 = -(s0 + tm1), rL1 = -fma(2.0, tm1, tp0), rL2 = fma(2.0, s2, tp0);
  const double uR = fma(oR2, rR2, fma(oR1, rR1, oR0 * rR0));
  const double uL = fma(oL2, rL2, fma(oL1, rL1, oL0 * rL0));
  const double hR0 = g_ur * oR0, hR1 = g_ur * oR1, hR2 = g_ur * oR2;
  const double hL0 = g_ul * oL0, hL1 = g_ul * oL1, hL2 = g_ul * oL2;
  // 1 / e_k from one reciprocal
  const double iP = fast_rcp(e0 * e12);
  const double G0 = -2.0 * (e12 * iP) * fma(hR0, rR0 - uR, hL0 * (rL0 - uL));
  const double G1 = -2.0 * (e02 * iP) * fma(hR1, rR1 - uR, hL1 * (rL1 - uL));
  const double G2 = -2.0 * (e01 * iP) * fma(hR2, rR2 - uR, hL2 * (rL2 - uL));
  // d e_k = 2 s_k d s_k + (26/3) dd_k d dd_k ; (26/3) dd_k recovered from p_k is not possible, so
  // the second differences are formed again (three subtractions)
  const double k263 = 26.0 / 3.0;
  const double da = tm1 - tm2, db = tp0 - tm1, dc = tp1 - tp0;
  const double A0 = G0 * s0, B0 = G0 * (k263 * da);
  const double A1 = G1 * s1, B1 = G1 * (k263 * db);
  const double A2 = G2 * s2, B2 = G2 * (k263 * dc);
  // cotangents of t(-2), t(-1), t(0), t(+1)
  double Tm2 = -(2.0 * A0 + B0);
  double Tm1 = fma(6.0, A0, B0) + fma(2.0, A1, -B1);
  double Tp0 = fma(2.0, A1, B1) - fma(6.0, A2, B2);
  double Tp1 = fma(2.0, A2, B2);
  Tm2 += fma(-2.0, hR0, hL0);
  Tm1 += fma(5.0, hR0, hR1) - fma(4.0, hL0, 2.0 * hL1);
  Tp0 += fma(2.0, hR1, 4.0 * hR2) - fma(5.0, hL2, hL1);
  Tp1 += fma(2.0, hL2, -hR2);
  const double k6 = 1.0 / 6.0;
  Weno5Vjp o;
  o.d[0] = -k6 * Tm2;
  o.d[1] = k6 * (Tm2 - Tm1);
  o.d[2] = fma(k6, Tm1 - Tp0, g_ur + g_ul);
  o.d[3] = k6 * (Tp0 - Tp1);
  o.d[4] = k6 * Tp1;
  return o;
}

// ESWENO32 (weno.py:284-296, reconstruction.py:413-439), the FAST form of esweno32_cell differentiated by hand:
// cotangents g_ur, g_ul of the face values and g_om of omega_0 (right-value weights: what the dissipative flux of
// the Burgers ESWENO32 scheme reads, burgers/schemes.py:237-247) -> cotangents of (m1, c, p1) in d[1..3].
//   d0 = c - m1, d1 = p1 - c, e_k = eps + d_k^2, tau = (d1 - d0)^2 (zero at the array ends),
//   A0 = (e0 + tau) e1, A1 = (e1 + tau) e0,
//   ur = c + (A0 d0 + 2 A1 d1) / (2 (A0 + 2 A1)),  ul = c - (A1 d1 + 2 A0 d0) / (2 (A1 + 2 A0)),  om0 = A0 / (A0 + 2 A1)
__device__ __forceinline__ Weno5Vjp esweno32_cell_vjp(double m1, double c, double p1, double eps, bool tau_zero,
                                                      double g_ur, double g_ul, double g_om) {
  const double d0 = c - m1, d1 = p1 - c;
  const double e0 = fma(d0, d0, eps), e1 = fma(d1, d1, eps);
  const double tt = d1 - d0;
  const double tau = tau_zero ? 0.0 : tt * tt;
  const double A0 = (e0 + tau) * e1, A1 = (e1 + tau) * e0;
  const double DR = fma(2.0, A1, A0), DL = fma(2.0, A0, A1);
  const double iR = 1.0 / DR, iL = 1.0 / DL;
  const double NR = fma(A0, d0, 2.0 * (A1 * d1)), NL = fma(A1, d1, 2.0 * (A0 * d0));
  // through the quotients
  const double gNR = 0.5 * g_ur * iR, gNL = -0.5 * g_ul * iL;
  const double gDR = -(gNR * NR + g_om * A0 * iR) * iR, gDL = -(gNL * NL) * iL;
  double gA0 = fma(gNR, d0, gDR) + g_om * iR + 2.0 * fma(gNL, d0, gDL);
  double gA1 = 2.0 * fma(gNR, d1, gDR) + fma(gNL, d1, gDL);
  double gd0 = gNR * A0 + 2.0 * (gNL * A0);
  double gd1 = 2.0 * (gNR * A1) + gNL * A1;
  // A0 = (e0 + tau) e1, A1 = (e1 + tau) e0
  const double ge0 = fma(gA0, e1, gA1 * (e1 + tau));
  const double ge1 = fma(gA0, e0 + tau, gA1 * e0);
  const double gtau = fma(gA0, e1, gA1 * e0);
  gd0 = fma(2.0 * d0, ge0, gd0);
  gd1 = fma(2.0 * d1, ge1, gd1);
  if (!tau_zero) {
    const double gtt = 2.0 * tt * gtau;
    gd1 += gtt;
    gd0 -= gtt;
  }
  Weno5Vjp o;
  o.d[0] = o.d[4] = 0.0;
  o.d[1] = -gd0;
  o.d[2] = (gd0 - gd1) + (g_ur + g_ul);
  o.d[3] = gd1;
  return o;
}

// partial derivatives of the dissipative flux g of esweno_gnum (unscaled) wrt (om_j, om_p, w_j, w_p)
struct EsGnumGrad {
  double d_omj, d_omp, d_wj, d_wp;
};
__device__ __forceinline__ EsGnumGrad esweno_gnum_grad(double omj, double omp, double wj, double wp, double delta) {
  const double dom = omp - omj;
  const double S = sqrt(fma(dom, dom, delta * delta));
  EsGnumGrad g;
  g.d_wp = -0.125 * (S + dom);
  g.d_wj = -g.d_wp;
  g.d_omp = -0.125 * (dom / S + 1.0) * (wp - wj);
  g.d_omj = -g.d_omp;
  return g;
}

template <int REC>
__device__ __forceinline__ Weno5Vjp reconstruct_cell_vjp(double m2, double m1, double c, double p1,
                                                         double p2, double eps, double g_ur,
                                                         double g_ul, bool tau_zero = false, double g_om = 0.0) {
  if (REC == PSK_REC_WENOJS53) return weno53_pair_vjp(m2, m1, c, p1, p2, eps, g_ur, g_ul);
  if (REC == PSK_REC_WENOJS32) return weno32_pair_vjp(m1, c, p1, eps, g_ur, g_ul);
  if (REC == PSK_REC_ESWENO32) return esweno32_cell_vjp(m1, c, p1, eps, tau_zero, g_ur, g_ul, g_om);
  Weno5Vjp o;
  o.d[0] = o.d[1] = o.d[3] = o.d[4] = 0.0;
  o.d[2] = g_ur + g_ul;
  return o;
}

// Partial derivatives of the face flux Phi(urj, ulp, wj, wp); d_speed is dPhi/d(global speed)
// for Lax-Friedrichs (scalar.py:277), zero otherwise.
struct FaceGrad {
  double d_ur, d_ul, d_wj, d_wp, d_speed;
};

__device__ __forceinline__ double sign0(double x) { return (x > 0.0) - (x < 0.0); }

template <int EQ, int FLUX>
__device__ __forceinline__ FaceGrad face_flux_grad(double urj, double ulp, double wj, double wp,
                                                    double speed, double nu, double arj,
                                                    double alp) {
  FaceGrad g;
  g.d_ur = g.d_ul = g.d_wj = g.d_wp = g.d_speed = 0.0;
  if (EQ == PSK_EQ_BURGERS) {
    if (FLUX == PSK_FLUX_RUSANOV || FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
      const double aj = fabs(wj), ap = fabs(wp);
      const double a = (FLUX == PSK_FLUX_LAX_FRIEDRICHS) ? speed : max_nan(ap, aj);
      const double half_an = 0.5 * a * nu;
      g.d_ur = 0.5 * urj + half_an;
      g.d_ul = 0.5 * ulp - half_an;
      const double d_a = -0.5 * nu * (ulp - urj);
      if (FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
        g.d_speed = d_a;
      } else {
        // jnp.maximum(|a[j+1]|, |a[j]|): larger argument takes the gradient, ties split
        const double sj = aj > ap ? 1.0 : (aj == ap ? 0.5 : 0.0);
        g.d_wj = d_a * sj * sign0(wj);
        g.d_wp = d_a * (1.0 - sj) * sign0(wp);
      }
    } else if (FLUX == PSK_FLUX_UPWIND || FLUX == PSK_FLUX_ESWENO) {  // (ESWENO: + esweno_gnum_grad)
      const bool pos = 0.5 * (urj + ulp) > 0.0;  // jnp.where: selected branch only
      g.d_ur = pos ? urj : 0.0;
      g.d_ul = pos ? 0.0 : ulp;
    } else {
      // f(max(ur, 0)) + f(min(ul, 0)); the tie at 0 carries f'(0) = 0
      g.d_ur = urj > 0.0 ? urj : 0.0;
      g.d_ul = ulp < 0.0 ? ulp : 0.0;
    }
    return g;
  }
  const bool pos = 0.5 * (arj + alp) > 0.0;
  if (EQ == PSK_EQ_ADVECTION) {
    g.d_ur = pos ? 1.0 : 0.0;
    g.d_ul = pos ? 0.0 : 1.0;
  } else {
    g.d_ur = pos ? arj : 0.0;
    g.d_ul = pos ? 0.0 : alp;
  }
  return g;
}

}  // namespace psk
