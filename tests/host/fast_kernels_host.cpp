// Host build (g++, no CUDA) of the specialised stage kernels, pyshocks_b200/csrc/psk_fast_kernels.cuh:
// the very device code of the product, run under the warp emulation of tests/host/emu/cuda_runtime.h
// (32 OS threads per warp, shuffles through barriers).  tests/test_fast_kernels_host.py checks
// every layout variant against the CPU oracle with it -- indexing, halo traffic, boundary
// handling, tails -- without a GPU.  Test infrastructure only.
#include <cuda_runtime.h>  // tests/host/emu/cuda_runtime.h (-I tests/host/emu)

#include <thread>
#include <vector>

#include "../../pyshocks_b200/csrc/psk_fast_kernels.cuh"

namespace emu {
thread_local emu_uint3 tid, bid, bdim, gdim;
thread_local Warp *warp = nullptr;
thread_local int lane = 0;
}  // namespace emu

namespace psk {
thread_local int g_last_cuda_error = 0;
}

namespace {

using Kernel = void (*)(const psk::FastParams);

// every (block, warp) of a (gx, gy) grid of CTAs of wpc warps, one warp at a time: 32 lane threads
// walk the same sequence and meet in the shuffles
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

// layout 0: stage_warp_fast_kernel, 2: stage_warp_fast_share_kernel with the u0 placement the product
// uses for the scheme, 12: the other placement (late u0 <-> early u0)
template <int EQ, int FLUX, int STAGE, bool WITH_MAX>
Kernel pick_layout(int layout, int /*late*/) {
  constexpr int kDefault = (EQ == PSK_EQ_BURGERS && FLUX == PSK_FLUX_RUSANOV) ? 0 : 1;
  if (layout == 2) return &psk::stage_warp_fast_share_kernel<EQ, FLUX, STAGE, WITH_MAX>;
  if (layout == 12) return &psk::stage_warp_fast_share_kernel<EQ, FLUX, STAGE, WITH_MAX, 1 - kDefault>;
  return &psk::stage_warp_fast_kernel<EQ, FLUX, STAGE, WITH_MAX>;
}

template <int EQ, int FLUX>
Kernel pick_stage(int stage, bool with_max, int layout, int late) {
  switch (stage * 2 + (with_max ? 1 : 0)) {
    case 0: return pick_layout<EQ, FLUX, 0, false>(layout, late);
    case 1: return pick_layout<EQ, FLUX, 0, true>(layout, late);
    case 2: return pick_layout<EQ, FLUX, 1, false>(layout, late);
    case 3: return pick_layout<EQ, FLUX, 1, true>(layout, late);
    case 4: return pick_layout<EQ, FLUX, 2, false>(layout, late);
    case 5: return pick_layout<EQ, FLUX, 2, true>(layout, late);
    case 6: return pick_layout<EQ, FLUX, 3, false>(layout, late);
    default: return pick_layout<EQ, FLUX, 3, true>(layout, late);
  }
}

template <int EQ, int FLUX>
double flux_scale() {
  return psk::FluxScale<EQ, FLUX>::value;
}

}  // namespace

extern "C" {

// One stage of the specialised kernel over `batch` rows, launched with the product's own geometry
// (psk::fast_geometry).  Returns 0, or -1 for a scheme the specialised kernels do not cover.
int emu_fast_stage(int layout, int late, int equation, int flux, int stage, int with_max, int bc, int n, int g,
                   int batch, long long ld, double dx, double eps, const double *uin, const double *u0,
                   double *uout, const double *dt, int dt_stride, const double *ghost, long long ghost_ld,
                   const double *lf_speed, unsigned long long *maxabs, const double *vel, const double *vel_l,
                   const double *vel_r, int wpc_max) {
  Kernel k = nullptr;
  double scale = 1.0;
  if (equation == PSK_EQ_BURGERS) {
    switch (flux) {
      case PSK_FLUX_RUSANOV:
        k = pick_stage<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV>(stage, with_max, layout, late);
        scale = flux_scale<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV>();
        break;
      case PSK_FLUX_LAX_FRIEDRICHS:
        k = pick_stage<PSK_EQ_BURGERS, PSK_FLUX_LAX_FRIEDRICHS>(stage, with_max, layout, late);
        scale = flux_scale<PSK_EQ_BURGERS, PSK_FLUX_LAX_FRIEDRICHS>();
        break;
      case PSK_FLUX_UPWIND:
        k = pick_stage<PSK_EQ_BURGERS, PSK_FLUX_UPWIND>(stage, with_max, layout, late);
        scale = flux_scale<PSK_EQ_BURGERS, PSK_FLUX_UPWIND>();
        break;
      case PSK_FLUX_ENGQUIST_OSHER:
        k = pick_stage<PSK_EQ_BURGERS, PSK_FLUX_ENGQUIST_OSHER>(stage, with_max, layout, late);
        scale = flux_scale<PSK_EQ_BURGERS, PSK_FLUX_ENGQUIST_OSHER>();
        break;
      default: return -1;
    }
  } else if (equation == PSK_EQ_ADVECTION && flux == PSK_FLUX_UPWIND) {
    k = pick_stage<PSK_EQ_ADVECTION, PSK_FLUX_UPWIND>(stage, with_max, layout, late);
  } else if (equation == PSK_EQ_CONTINUITY && flux == PSK_FLUX_UPWIND) {
    k = pick_stage<PSK_EQ_CONTINUITY, PSK_FLUX_UPWIND>(stage, with_max, layout, late);
  } else {
    return -1;
  }

  psk::FastParams q{};
  q.uin = uin; q.u0 = u0; q.uout = uout; q.dt = dt; q.lf_speed = lf_speed;
  q.maxabs = with_max ? maxabs : nullptr;
  q.vel = vel; q.vel_l = vel_l; q.vel_r = vel_r;
  q.bc.ghost = ghost; q.bc.ghost_ld = ghost_ld; q.bc.bc = bc; q.bc.n = n; q.bc.g = g; q.bc.nx = n + 2 * g;
  q.ld = ld;
  q.coef = (1.0 / dx) / scale;
  q.eps9 = eps * (1.0 / 9.0);
  q.dt_stride = dt_stride;
  const psk::FastGeometry geo = psk::fast_geometry(n, wpc_max);
  q.chunks_per_row = geo.chunks_per_row;
  const unsigned gx = static_cast<unsigned>((geo.chunks_per_row + geo.wpc - 1) / geo.wpc);

  run_grid(gx, static_cast<unsigned>(batch), geo.wpc, [&]() { k(q); });
  return 0;
}

// One whole SSPRK33 step with the fused kernel (R cells per lane), launched like launch_step_shape.
int emu_fused_step(int R, int flux, int bc_none, int with_max, int n, int g, int batch, long long ld, double dx, double eps,
                   const double *u, double *uout, const double *dt, int dt_stride, const unsigned char *active,
                   unsigned long long *maxabs) {
  psk::StepParams q{};
  q.u = u; q.uout = uout; q.dt = dt; q.active = active;
  q.maxabs = with_max ? maxabs : nullptr;
  q.ld = ld;
  q.coef = (1.0 / dx) / (flux == PSK_FLUX_RUSANOV ? 4.0 : 2.0);
  q.eps9 = eps * (1.0 / 9.0);
  q.dt_stride = dt_stride;
  q.n = n;
  q.g = g;
  q.bc_none = bc_none;
  void (*k)(const psk::StepParams) = nullptr;
  int emit = 0;
  constexpr int kRus = PSK_FLUX_RUSANOV, kUp = PSK_FLUX_UPWIND, kEo = PSK_FLUX_ENGQUIST_OSHER;
#define EMU_STEP(RR, FL)                                                                                    \
  do {                                                                                                      \
    k = with_max ? &psk::step_warp_fused_kernel<RR, FL, true, 256, 1> : &psk::step_warp_fused_kernel<RR, FL, false, 256, 1>; \
    emit = psk::StepGeometry<RR>::kEmit;                                                                    \
  } while (0)
  if (flux == kRus && R == 4) EMU_STEP(4, kRus);
  else if (flux == kRus && R == 6) EMU_STEP(6, kRus);
  else if (flux == kRus && R == 8) EMU_STEP(8, kRus);
  else if (flux == kRus && R == 10) EMU_STEP(10, kRus);
  else if (flux == kUp && R == 6) EMU_STEP(6, kUp);
  else if (flux == kEo && R == 6) EMU_STEP(6, kEo);
  else return -1;
#undef EMU_STEP
  q.chunks_per_row = (n + emit - 1) / emit;
  int wpc = 8;
  if (q.chunks_per_row < wpc) wpc = q.chunks_per_row;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  run_grid(gx, static_cast<unsigned>(batch), wpc, [&]() { k(q); });
  return 0;
}

// The STAGES form of the whole-step kernel (psk_ssprk33_step_stages): k1, k2 stored, uout optional.
int emu_fused_step_stages(int n, int g, int batch, long long ld, double dx, double eps, const double *u, double *k1,
                          double *k2, double *uout, const double *dt, int dt_stride) {
  psk::StepParams q{};
  q.u = u; q.uout = uout; q.dt = dt; q.k1_out = k1; q.k2_out = k2;
  q.ld = ld;
  q.coef = (1.0 / dx) / 4.0;
  q.eps9 = eps * (1.0 / 9.0);
  q.dt_stride = dt_stride;
  q.n = n;
  q.g = g;
  q.chunks_per_row = (n + psk::StepGeometry<6>::kEmit - 1) / psk::StepGeometry<6>::kEmit;
  int wpc = 4;
  if (q.chunks_per_row < wpc) wpc = q.chunks_per_row;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  run_grid(gx, static_cast<unsigned>(batch), wpc,
           [&]() { psk::step_warp_fused_kernel<6, PSK_FLUX_RUSANOV, false, 128, 3, true>(q); });
  return 0;
}

// psk_ssprk33_step_bc: the whole-step kernel on rows with Dirichlet data at the three stage times (ghost3: three
// blocks of batch * ghost_ld or 2 g doubles), Burgers fluxes or the advection / continuity upwind flux.
int emu_fused_step_bc(int equation, int flux, int with_max, int neumann, int n, int g, int batch, long long ld, double dx, double eps,
                      const double *u, double *uout, const double *dt, int dt_stride, const double *ghost3,
                      long long ghost_ld, const double *vel, const double *vel_l, const double *vel_r,
                      unsigned long long *maxabs, const double *nu) {
  psk::StepParams q{};
  q.nu = nu;
  q.u = u; q.uout = uout; q.dt = dt;
  q.maxabs = with_max ? maxabs : nullptr;
  q.ld = ld;
  q.coef = (1.0 / dx) / (equation != PSK_EQ_BURGERS ? 1.0 : (flux == PSK_FLUX_RUSANOV ? 4.0 : 2.0));
  q.eps9 = eps * (1.0 / 9.0);
  q.dt_stride = dt_stride;
  q.n = n;
  q.g = g;
  q.ghost3 = ghost3;
  q.ghost_ld = ghost_ld;
  q.ghost_block = ghost_ld != 0 ? static_cast<long long>(batch) * ghost_ld : 2 * g;
  q.vel = vel; q.vel_l = vel_l; q.vel_r = vel_r;
  void (*k)(const psk::StepParams) = nullptr;
  constexpr int kB = PSK_EQ_BURGERS, kUp = PSK_FLUX_UPWIND;
#define EMU_BC(FL, EQ)                                                                                          \
  k = neumann ? (with_max ? &psk::step_warp_fused_kernel<6, FL, true, 128, 3, false, EQ, 2>                     \
                          : &psk::step_warp_fused_kernel<6, FL, false, 128, 3, false, EQ, 2>)                   \
              : (with_max ? &psk::step_warp_fused_kernel<6, FL, true, 128, 3, false, EQ, 1>                     \
                          : &psk::step_warp_fused_kernel<6, FL, false, 128, 3, false, EQ, 1>)
  if (equation == PSK_EQ_ADVECTION && flux == kUp) { EMU_BC(kUp, PSK_EQ_ADVECTION); }
  else if (equation == PSK_EQ_CONTINUITY && flux == kUp) { EMU_BC(kUp, PSK_EQ_CONTINUITY); }
  else if (equation == kB && flux == PSK_FLUX_RUSANOV && nu != nullptr) {  // alpha != 1: nu of every face
    k = neumann ? &psk::step_warp_fused_kernel<6, PSK_FLUX_RUSANOV, false, 128, 3, false, kB, 2, true>
                : &psk::step_warp_fused_kernel<6, PSK_FLUX_RUSANOV, false, 128, 3, false, kB, 1, true>;
  }
  else if (equation == kB && flux == PSK_FLUX_RUSANOV) { EMU_BC(PSK_FLUX_RUSANOV, kB); }
  else if (equation == kB && flux == kUp) { EMU_BC(kUp, kB); }
  else if (equation == kB && flux == PSK_FLUX_ENGQUIST_OSHER) { EMU_BC(PSK_FLUX_ENGQUIST_OSHER, kB); }
  else return -1;
#undef EMU_BC
  q.chunks_per_row = (n + psk::StepGeometry<6>::kEmit - 1) / psk::StepGeometry<6>::kEmit;
  int wpc = 4;
  if (q.chunks_per_row < wpc) wpc = q.chunks_per_row;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  run_grid(gx, static_cast<unsigned>(batch), wpc, [&]() { k(q); });
  return 0;
}

// the whole-step kernel for the advection / continuity equations on PERIODIC rows (psk_ssprk33_step)
int emu_fused_step_periodic_eq(int equation, int n, int g, int batch, long long ld, double dx, double eps, const double *u,
                               double *uout, const double *dt, int dt_stride, const double *vel, const double *vel_l,
                               const double *vel_r) {
  psk::StepParams q{};
  q.u = u; q.uout = uout; q.dt = dt;
  q.ld = ld;
  q.coef = 1.0 / dx;
  q.eps9 = eps * (1.0 / 9.0);
  q.dt_stride = dt_stride;
  q.n = n;
  q.g = g;
  q.vel = vel; q.vel_l = vel_l; q.vel_r = vel_r;
  void (*k)(const psk::StepParams) = nullptr;
  if (equation == PSK_EQ_ADVECTION) k = &psk::step_warp_fused_kernel<6, PSK_FLUX_UPWIND, false, 128, 3, false, PSK_EQ_ADVECTION, 0>;
  else if (equation == PSK_EQ_CONTINUITY) k = &psk::step_warp_fused_kernel<6, PSK_FLUX_UPWIND, false, 128, 3, false, PSK_EQ_CONTINUITY, 0>;
  else return -1;
  q.chunks_per_row = (n + psk::StepGeometry<6>::kEmit - 1) / psk::StepGeometry<6>::kEmit;
  int wpc = 4;
  if (q.chunks_per_row < wpc) wpc = q.chunks_per_row;
  const unsigned gx = static_cast<unsigned>((q.chunks_per_row + wpc - 1) / wpc);
  run_grid(gx, static_cast<unsigned>(batch), wpc, [&]() { k(q); });
  return 0;
}

int emu_chunks_per_row(int n) { return psk::fast_geometry(n, 8).chunks_per_row; }

}  // extern "C"
