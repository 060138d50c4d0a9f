// psk_adjoint_math.cuh -- per-cell arithmetic of the lean adjoint stage kernel (fp64).
//
// Same derivative as weno53_pair_vjp_sixths (psk_math.cuh), re-derived on the PRODUCT form of
// the WENO-JS weights that the forward kernels evaluate (weno.py:247-256 after multiplying
// numerator and denominator by (e0 e1 e2)^2):
//
//   U_R - c = N_R / D_R,  N_R = sum_k a_k r_k,  D_R = sum_k a_k,
//   (a_0, a_1, a_2) = (W0, 6 W1, 3 W2) for the right value, (3 W0, 6 W1, W2) for the left one,
//   W0 = (e1 e2)^2, W1 = (e0 e2)^2, W2 = (e0 e1)^2,  e_k = s_k^2 + (13/3) dd_k^2 + eps / 9
//
// (everything in SIXTHS of first differences, see weno53_pair_lean).  Differentiating the
// products instead of the quotients d_k / e_k^2 needs no 1 / e_k at all:
//
//   z = g / D                      cotangent of N;  cot(a_k) = z (r_k - (U - c))
//   cot(E12) = 2 E12 cot(W0), ...  cot(e_0) = e2 cot(E02) + e1 cot(E01), ...
//
// and the cotangents are accumulated on the four first differences t(-2..1) of the cell
// (shared by neighbouring cells), not on the five cell values: the conversion to cells,
// cot(u_j) = (T(j-1, j) - T(j, j+1)) / 6, is done once per cell at the very end.
//
// The functions are plain `double` arithmetic (fma, *, +) so that tests/test_adjoint_math.py
// can compile this header for the HOST (g++) and check the derivative against central
// differences of the forward value without a GPU.
#pragma once

#ifndef PSK_HD
#define PSK_HD __device__ __forceinline__
#endif

namespace psk {

// what the forward pass of a cell leaves behind for its vector-Jacobian product
struct Weno5State {
  double iR, iL;  // 1 / D_R, 1 / D_L
  double uR, uL;  // U_R - c, U_L - c
  double e0, e1, e2;
  double E12, E02, E01;
  double W0, a1, W2;  // a1 = 6 W1
};

// t(-2), t(-1), t(0), t(+1): (u[k+1] - u[k]) / 6 around the cell; pm1, p0, pp1 = (13/3) dd^2 + eps/9
// for the second differences centred at i-1, i, i+1.
PSK_HD Weno5State weno53_state(double tm2, double tm1, double tp0, double tp1, double pm1, double p0,
                               double pp1) {
  const double s0 = fma(3.0, tm1, -tm2);
  const double s1 = tp0 + tm1;
  const double s2 = fma(-3.0, tp0, tp1);
  Weno5State F;
  F.e0 = fma(s0, s0, pm1);
  F.e1 = fma(s1, s1, p0);
  F.e2 = fma(s2, s2, pp1);
  F.E12 = F.e1 * F.e2;
  F.E02 = F.e0 * F.e2;
  F.E01 = F.e0 * F.e1;
  F.W0 = F.E12 * F.E12;
  F.W2 = F.E01 * F.E01;
  F.a1 = 6.0 * (F.E02 * F.E02);
  const double a2R = 3.0 * F.W2, a0L = 3.0 * F.W0;
  const double rR0 = fma(2.0, s0, -tm1), rR1 = fma(2.0, tp0, tm1), rR2 = tp0 - s2;
  const double nL0 = s0 + tm1, nL1 = fma(2.0, tm1, tp0), rL2 = fma(2.0, s2, tp0);
  const double numR = fma(a2R, rR2, fma(F.a1, rR1, F.W0 * rR0));
  const double numL = fma(F.W2, rL2, -fma(F.a1, nL1, a0L * nL0));
  F.iR = fast_rcp((F.W0 + F.a1) + a2R);
  F.iL = fast_rcp((a0L + F.a1) + F.W2);
  F.uR = numR * F.iR;
  F.uL = numL * F.iL;
  return F;
}

// Adds the cell's contribution to the cotangents T(-2), T(-1), T(0), T(+1) of its four first
// differences; gR / gL are the cotangents of U_R / U_L.  (The direct term dU/dc = 1 is the
// caller's: cot(c) += gR + gL.)
PSK_HD void weno53_vjp_acc(const Weno5State &F, double tm2, double tm1, double tp0, double tp1,
                           double gR, double gL, double &Tm2, double &Tm1, double &Tp0, double &Tp1) {
  const double s0 = fma(3.0, tm1, -tm2);
  const double s1 = tp0 + tm1;
  const double s2 = fma(-3.0, tp0, tp1);
  const double rR0 = fma(2.0, s0, -tm1), rR1 = fma(2.0, tp0, tm1), rR2 = tp0 - s2;
  const double nL0 = s0 + tm1, nL1 = fma(2.0, tm1, tp0), rL2 = fma(2.0, s2, tp0);
  const double zR = gR * F.iR, zL = gL * F.iL;
  const double zR3 = 3.0 * zR, zL3 = 3.0 * zL;
  // r_k - (U - c): deviations only, no cancellation against the cell value
  const double dR0 = rR0 - F.uR, dR1 = rR1 - F.uR, dR2 = rR2 - F.uR;
  const double mL0 = nL0 + F.uL, mL1 = nL1 + F.uL, dL2 = rL2 - F.uL;  // mL = -(rL - uL)
  const double cW0 = fma(zR, dR0, -(zL3 * mL0));
  const double cW1 = fma(zR, dR1, -(zL * mL1));  // times 6
  const double cW2 = fma(zR3, dR2, zL * dL2);
  const double X12 = F.E12 * cW0;
  const double Y02 = 6.0 * (F.E02 * cW1);
  const double X01 = F.E01 * cW2;
  // cot(e_k) / 2
  const double G0 = fma(F.e2, Y02, F.e1 * X01);
  const double G1 = fma(F.e2, X12, F.e0 * X01);
  const double G2 = fma(F.e1, X12, F.e0 * Y02);
  // e_k = s_k^2 + (13/3) dd_k^2 + eps/9:  cot(s_k) = 4 s_k G_k,  cot(dd_k) = (52/3) dd_k G_k
  const double A0 = s0 * G0, A1 = s1 * G1, A2 = s2 * G2;
  const double B0 = (tm1 - tm2) * G0, B1 = (tp0 - tm1) * G1, B2 = (tp1 - tp0) * G2;
  // cotangents of the candidate offsets r_k
  const double hR0 = zR * F.W0, hR1 = zR * F.a1, hR2 = zR3 * F.W2;
  const double hL0 = zL3 * F.W0, hL1 = zL * F.a1, hL2 = zL * F.W2;
  // s0 and s2 also appear in rR0 = 2 s0 - t(-1), rL0 = -(s0 + t(-1)), rR2 = t(0) - s2, rL2 = 2 s2 + t(0)
  const double S0 = fma(4.0, A0, fma(2.0, hR0, -hL0));
  const double S2 = fma(4.0, A2, fma(2.0, hL2, -hR2));
  const double k = 52.0 / 3.0;
  const double A14 = 4.0 * A1;
  Tm2 = fma(-k, B0, Tm2) - S0;
  Tp1 = fma(k, B2, Tp1) + S2;
  // t(-1): 3 S0 + 4 A1 + k (B0 - B1) - hR0 + hR1 - hL0 - 2 hL1
  const double y1 = (hR1 - hR0) - fma(2.0, hL1, hL0);
  Tm1 = fma(k, B0 - B1, fma(3.0, S0, Tm1)) + (A14 + y1);
  // t(0): -3 S2 + 4 A1 + k (B1 - B2) + 2 hR1 + hR2 - hL1 + hL2
  const double y0 = fma(2.0, hR1, hR2) + (hL2 - hL1);
  Tp0 = fma(k, B1 - B2, fma(-3.0, S2, Tp0)) + (A14 + y0);
}

}  // namespace psk
