// Host build of pyshocks_b200/csrc/psk_adjoint_math.cuh (g++, no CUDA): the lean adjoint
// arithmetic applied cell by cell to one array, for tests/test_adjoint_math.py.
// Test infrastructure only.
#include <cmath>
#include <vector>

#define PSK_HD inline
namespace psk {
inline double fast_rcp(double x) { return 1.0 / x; }  // the device version is a ~1 ulp reciprocal
}  // namespace psk
using std::fma;
#include "../../pyshocks_b200/csrc/psk_adjoint_math.cuh"

extern "C" {

// f[n] -> fl[n], fr[n] (left / right face values) for cells 2 .. n-3, and the cotangent
// d[n] = (d fl / d f)^T gl + (d fr / d f)^T gr restricted to those cells' stencils.
void lean_reconstruct_vjp(int n, const double *f, double eps, const double *gl, const double *gr,
                          double *fl, double *fr, double *d) {
  std::vector<double> t(n, 0.0), T(n, 0.0), pq(n, 0.0);
  const double eps9 = eps / 9.0;
  for (int k = 0; k + 1 < n; ++k) t[k] = (1.0 / 6.0) * (f[k + 1] - f[k]);  // t[k]: cells (k, k+1)
  for (int k = 0; k + 2 < n; ++k) {  // pq[k]: second difference centred at cell k + 1
    const double dd = t[k + 1] - t[k];
    pq[k] = fma((13.0 / 3.0) * dd, dd, eps9);
  }
  for (int i = 0; i < n; ++i) d[i] = 0.0;
  for (int i = 2; i + 2 < n; ++i) {
    const psk::Weno5State F = psk::weno53_state(t[i - 2], t[i - 1], t[i], t[i + 1], pq[i - 2], pq[i - 1], pq[i]);
    fr[i] = f[i] + F.uR;
    fl[i] = f[i] + F.uL;
    psk::weno53_vjp_acc(F, t[i - 2], t[i - 1], t[i], t[i + 1], gr[i], gl[i], T[i - 2], T[i - 1], T[i], T[i + 1]);
    d[i] += gr[i] + gl[i];
  }
  // cot(f_j) = (T(j-1, j) - T(j, j+1)) / 6
  for (int j = 0; j < n; ++j) {
    const double a = j >= 1 ? T[j - 1] : 0.0, b = j + 1 < n ? T[j] : 0.0;
    d[j] += (a - b) / 6.0;
  }
}

}  // extern "C"
