// Host build (g++, no CUDA) of the fused reverse SSPRK33 step, pyshocks_b200/csrc/psk_reverse_kernels.cuh:
// the device code of the product under the warp emulation of tests/host/emu/cuda_runtime.h.
// tests/test_reverse_kernel_host.py checks it against reverse-mode differentiation of the reference
// arithmetic (oracle/torch_twin.py) and its recomputed stage values against the C oracle.
// Test infrastructure only.
#include <cuda_runtime.h>  // tests/host/emu/cuda_runtime.h (-I tests/host/emu)

#include <thread>
#include <vector>

#include "../../pyshocks_b200/csrc/psk_reverse_kernels.cuh"

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
void run_grid(unsigned gx, unsigned gy, Body body) {
  emu::Warp warp;
  pthread_barrier_init(&warp.bar, nullptr, 32);
  std::vector<std::thread> lanes;
  for (int lane = 0; lane < 32; ++lane) {
    lanes.emplace_back([&, lane]() {
      emu::warp = &warp;
      emu::lane = lane;
      emu::bdim = {32u, 1u, 1u};
      emu::gdim = {gx, gy, 1u};
      for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) {
          emu::bid = {bx, by, 0u};
          emu::tid = {static_cast<unsigned>(lane), 0u, 0u};
          body();
        }
    });
  }
  for (auto &t : lanes) t.join();
  pthread_barrier_destroy(&warp.bar);
}

template <int C>
void run(psk::RevParams p, int batch) {
  psk::rev_tiling(p.n, C, p.tiles_per_row, p.c_last);
  run_grid(static_cast<unsigned>(p.tiles_per_row), static_cast<unsigned>(batch), [&]() {
    if (p.ghost3 != nullptr) psk::reverse_step_kernel<C, 1, true>(p);
    else psk::reverse_step_kernel<C, 1>(p);
  });
}

}  // namespace

extern "C" {

// p_out = (d advance / d u)^T p_in on periodic rows (interior cells) or, with ghost3, Dirichlet rows; launched like
// launch_reverse<C>
int emu_reverse_step(int C, int n, int g, int batch, long long ld, double dx, double eps, const double *u,
                     const double *pin, const double *dt, int dt_stride, double *pout, double *k1, double *k2,
                     int bc_none, const double *ghost3, long long ghost_ld, long long ghost_block) {
  psk::RevParams p{};
  p.bc_none = bc_none;
  p.ghost3 = ghost3; p.ghost_ld = ghost_ld; p.ghost_block = ghost_block;  // Dirichlet rows (ghost3 != NULL)
  p.u = u; p.pin = pin; p.pout = pout; p.dt = dt; p.dt_stride = dt_stride; p.ld = ld;
  p.invdx = 1.0 / dx; p.eps = eps; p.n = n; p.g = g;
  p.dbg_k1 = k1; p.dbg_k2 = k2;
  switch (C) {
    case 8: run<8>(p, batch); break;
    case 12: run<12>(p, batch); break;
    case 16: run<16>(p, batch); break;
    case 20: run<20>(p, batch); break;
    case 24: run<24>(p, batch); break;
    default: return -1;
  }
  return 0;
}

}  // extern "C"
