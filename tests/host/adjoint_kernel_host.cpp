// Host build (g++, no CUDA) of the lean adjoint stage kernel, pyshocks_b200/csrc/psk_adjoint_kernels.cuh:
// the device code of the product under the warp emulation of tests/host/emu/cuda_runtime.h, followed
// by the transpose of the periodic boundary fill (what adjoint_boundary_kernel does on the GPU).
// tests/test_adjoint_kernel_host.py checks it against reverse-mode differentiation of the reference
// arithmetic (oracle/torch_twin.py).  Test infrastructure only.
#include <cuda_runtime.h>  // tests/host/emu/cuda_runtime.h (-I tests/host/emu)

#include <thread>
#include <vector>

#include "../../pyshocks_b200/csrc/psk_adjoint_kernels.cuh"

namespace emu {
thread_local emu_uint3 tid, bid, bdim, gdim;
thread_local Warp *warp = nullptr;
thread_local int lane = 0;
}  // namespace emu

namespace psk {
thread_local int g_last_cuda_error = 0;
}

namespace {

template <class Body>
void run_grid(unsigned gx, unsigned gy, int wpc, Body body) {
  emu::Warp warp;
  pthread_barrier_init(&warp.bar, nullptr, 32);
  std::vector<std::thread> lanes;
  for (int lane = 0; lane < 32; ++lane) {
    lanes.emplace_back([&, lane]() {
      emu::warp = &warp;
      emu::lane = lane;
      emu::bdim = {static_cast<unsigned>(wpc * 32), 1u, 1u};
      emu::gdim = {gx, gy, 1u};
      for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx)
          for (int wi = 0; wi < wpc; ++wi) {
            emu::bid = {bx, by, 0u};
            emu::tid = {static_cast<unsigned>(wi * 32 + lane), 0u, 0u};
            body();
          }
    });
  }
  for (auto &t : lanes) t.join();
  pthread_barrier_destroy(&warp.bar);
}

}  // namespace

extern "C" {

// out = c_acc acc + c_acc2 acc2 + c_v (v + c_g dt J_L(x)^T v) on periodic rows (Burgers, Rusanov,
// WENO-JS5), launched like launch_adjoint_lean; variant: bit 0 = linear part parked in shared
// memory, bit 1 = state of cell 3 recomputed (3 is the product's default).
int emu_adjoint_lean(int variant, int n, int g, int batch, long long ld, double dx, double eps, const double *x,
                     const double *v, const double *dt, int dt_stride, double c_v, double c_g, const double *acc,
                     double c_acc, const double *acc2, double c_acc2, double *gspill, double *out) {
  psk::AdjParams p{};
  p.x = x; p.v = v; p.acc = acc; p.acc2 = acc2; p.out = out; p.dt = dt; p.dt_stride = dt_stride;
  p.c_v = c_v; p.c_g = c_g; p.c_acc = c_acc; p.c_acc2 = c_acc2;
  p.gspill = gspill;
  p.bc.ghost = nullptr; p.bc.ghost_ld = 0; p.bc.bc = PSK_BC_PERIODIC; p.bc.n = n; p.bc.g = g; p.bc.nx = n + 2 * g;
  p.ld = ld;
  p.invdx = 1.0 / dx;
  p.eps = eps;
  p.prescaled = 1;
  const int chunks = (n + g + 119) / 120;
  const int total = chunks + 1;
  const int wpc = total < 4 ? total : 4;
  const unsigned gx = static_cast<unsigned>((total + wpc - 1) / wpc);
  void (*k)(const psk::AdjParams, int) = nullptr;
  switch (variant) {
    case 0: k = &psk::adjoint_lean_kernel<4, 0>; break;
    case 3: k = &psk::adjoint_lean_kernel<4, 3>; break;
    default: return -1;
  }
  run_grid(gx, static_cast<unsigned>(batch), wpc, [&]() { k(p, chunks); });
  // transpose of the periodic boundary fill (adjoint_boundary_kernel; gspill is pre-scaled by c_g dt)
  const int nx = n + 2 * g;
  for (int row = 0; row < batch; ++row)
    for (int kk = 0; kk < 2 * g; ++kk) {
      const int i = kk < g ? kk : nx - 2 * g + kk;
      const int src = i < g ? i + n : i - n;
      out[static_cast<long long>(row) * ld + src] += gspill[static_cast<long long>(row) * 2 * g + kk];
    }
  return 0;
}

// The Lax-Friedrichs (flux = 1) and alpha != 1 (nu != NULL: the viscosity of every face) forms of the lean kernel, on
// periodic (bck = 0), Dirichlet (bck = 1, ghost: 2 g data per row) or Neumann (bck = 2, ghost: 2 g offsets per row) rows,
// followed by what adjoint_boundary_kernel
// does on the GPU: the transpose of the periodic fill and, for Lax-Friedrichs, the cotangent of the row's speed
// max |w| shared equally between the arg-max cells (ghost cells pass theirs on to the cells they copy; Dirichlet
// data drop it).
int emu_adjoint_lean_flux(int flux, int bck, int n, int g, int batch, long long ld, double dx, double eps, const double *x,
                          const double *v, const double *dt, int dt_stride, double c_v, double c_g, const double *nu,
                          const double *ghost, double *gspill, double *out) {
  psk::AdjParams p{};
  p.x = x; p.v = v; p.out = out; p.dt = dt; p.dt_stride = dt_stride;
  p.c_v = c_v; p.c_g = c_g;
  p.gspill = gspill;
  p.nu = nu;
  p.bc.ghost = ghost; p.bc.ghost_ld = ghost != nullptr ? 2 * g : 0;
  p.bc.bc = bck == 0 ? PSK_BC_PERIODIC : (bck == 1 ? PSK_BC_DIRICHLET : PSK_BC_NEUMANN); p.bc.n = n; p.bc.g = g; p.bc.nx = n + 2 * g;
  p.ld = ld;
  p.invdx = 1.0 / dx;
  p.eps = eps;
  p.prescaled = 1;
  const int nx = n + 2 * g;
  std::vector<double> speed(batch, 0.0), ga(batch, 0.0);
  for (int row = 0; row < batch; ++row)
    for (int i = 0; i < nx; ++i) speed[row] = std::fmax(speed[row], std::fabs(psk::load_w(p.bc, x + row * ld, row, i)));
  p.speed = speed.data();
  p.ga = ga.data();
  std::vector<unsigned> amax(2 * static_cast<size_t>(batch), 0u);
  p.amax = amax.data();
  const int chunks = (n + g + 119) / 120;
  const int total = chunks + 1;
  const int wpc = total < 4 ? total : 4;
  const unsigned gx = static_cast<unsigned>((total + wpc - 1) / wpc);
  void (*k)(const psk::AdjParams, int) = nullptr;
  if (flux == 1) k = nu != nullptr ? &psk::adjoint_lean_kernel<4, 3, PSK_FLUX_LAX_FRIEDRICHS, true>
                                   : &psk::adjoint_lean_kernel<4, 3, PSK_FLUX_LAX_FRIEDRICHS, false>;
  else k = nu != nullptr ? &psk::adjoint_lean_kernel<4, 3, PSK_FLUX_RUSANOV, true> : &psk::adjoint_lean_kernel<4, 3>;
  run_grid(gx, static_cast<unsigned>(batch), wpc, [&]() { k(p, chunks); });
  // where apply_boundary copied ghost cell i from (periodic image, Neumann mirror image; Dirichlet data: from nowhere)
  auto source = [&](int i) {
    if (i >= g && i < nx - g) return i;
    if (bck == 0) return i < g ? i + n : i - n;
    if (bck == 2) return i < g ? 2 * g - 1 - i : 2 * (nx - g) - 1 - i;
    return -1;
  };
  for (int row = 0; row < batch; ++row) {
    double *orow = out + static_cast<long long>(row) * ld;
    if (bck != 1)
      for (int kk = 0; kk < 2 * g; ++kk) orow[source(kk < g ? kk : nx - 2 * g + kk)] += gspill[static_cast<long long>(row) * 2 * g + kk];
    if (flux == 1 && amax[2 * row] == 1u) {  // a single arg-max cell: the index the kernel recorded, no scan
      const int i = static_cast<int>(amax[2 * row + 1]);
      const double wi = psk::load_w(p.bc, x + row * ld, row, i);
      if (std::fabs(wi) != speed[row]) return -2;
      if (source(i) >= 0) orow[source(i)] += ga[row] * (wi > 0.0 ? 1.0 : (wi < 0.0 ? -1.0 : 0.0));
    } else if (flux == 1) {
      int count = 0;
      for (int i = 0; i < nx; ++i) count += std::fabs(psk::load_w(p.bc, x + row * ld, row, i)) == speed[row];
      if (static_cast<unsigned>(count) != amax[2 * row]) return -3;  // the kernel counted every arg-max cell of the array once
      for (int i = 0; i < nx; ++i) {
        const double wi = psk::load_w(p.bc, x + row * ld, row, i);
        if (std::fabs(wi) == speed[row] && source(i) >= 0) orow[source(i)] += ga[row] / count * (wi > 0.0 ? 1.0 : (wi < 0.0 ? -1.0 : 0.0));
      }
    }
  }
  return 0;
}

}  // extern "C"
