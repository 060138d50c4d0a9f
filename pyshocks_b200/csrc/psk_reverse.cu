// psk_reverse.cu -- launch of the fused reverse SSPRK33 step (psk_reverse_kernels.cuh) behind the C ABI.
#include <cuda_runtime.h>

#include <cstdint>

#include "psk_common.cuh"
#include "psk_reverse_kernels.cuh"

namespace psk {

int g_reverse_variant = 0;  // 0: pick the run length per row length; 12 / 16 / 20 / 24: forced

// warps (CTAs) resident per SM by shared memory: 3 window arrays + the parked values + 1 KB per CTA
template <int C>
constexpr int rev_min_blocks() {
  constexpr int bytes = RevGeometry<C>::kSmemDoubles * 8 + 1024;
  constexpr int by_smem = (227 * 1024) / bytes;
  return by_smem > 12 ? 12 : by_smem;  // 168 registers per thread: no spills
}

template <int C>
int launch_reverse(const RevParams &p0, int batch, cudaStream_t st) {
  RevParams p = p0;
  rev_tiling(p.n, C, p.tiles_per_row, p.c_last);
  unsigned gy, gz;
  if (!split_rows(batch, gy, gz)) return PSK_E_UNSUPPORTED;
  const dim3 grid(static_cast<unsigned>(p.tiles_per_row), gy, gz);
  if (p.ghost3 != nullptr)
    reverse_step_kernel<C, rev_min_blocks<C>(), true><<<grid, 32, 0, st>>>(p);
  else
    reverse_step_kernel<C, rev_min_blocks<C>()><<<grid, 32, 0, st>>>(p);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

// shared body of psk_ssprk33_step_adjoint / psk_ssprk33_step_adjoint_bc
static int run_reverse(const psk_desc *d, const double *u, const double *p_in, const double *dt, int64_t dt_stride,
                       const double *ghost3, double *p_out, double *k1_out, double *k2_out, cudaStream_t st) {
  int rc = check_desc(d);
  if (rc != PSK_OK) return rc;
  if (u == nullptr || p_in == nullptr || p_out == nullptr || dt == nullptr) return PSK_E_INVALID;
  if (p_out == p_in || p_out == u || (k1_out == nullptr) != (k2_out == nullptr)) return PSK_E_INVALID;
  auto al = [&](const double *a) { return a == nullptr || reinterpret_cast<uintptr_t>(a + d->g) % 16 == 0; };
  const bool aligned = al(u) && al(p_in) && al(p_out) && al(k1_out) && al(k2_out) && (d->ld % 2 == 0);
  const bool dirichlet = d->bc == PSK_BC_DIRICHLET;
  if (dirichlet && ghost3 == nullptr && d->ghost == nullptr) return PSK_E_INVALID;
  if (d->equation != PSK_EQ_BURGERS || d->flux != PSK_FLUX_RUSANOV || d->rec != PSK_REC_WENOJS53 ||
      d->math != PSK_MATH_FAST || d->nu != nullptr || !aligned || d->n % 2 != 0 || d->n < 8 ||
      !((d->bc == PSK_BC_PERIODIC && d->g >= 3) || (d->bc == PSK_BC_NONE && d->g >= kRevHalo) ||
        (dirichlet && d->g == 3)))
    return PSK_E_UNSUPPORTED;
  RevParams p{};
  p.u = u; p.pin = p_in; p.pout = p_out; p.dt = dt; p.dt_stride = dt_stride; p.ld = d->ld;
  p.invdx = 1.0 / d->dx;
  p.eps = d->eps;
  p.n = d->n; p.g = d->g;
  p.dbg_k1 = k1_out; p.dbg_k2 = k2_out;
  p.bc_none = d->bc == PSK_BC_NONE ? 1 : 0;
  if (dirichlet) {  // data of the three stage times, or the descriptor's time-independent data for all of them
    p.ghost3 = ghost3 != nullptr ? ghost3 : d->ghost;
    p.ghost_ld = d->ghost_ld;
    p.ghost_block = ghost3 != nullptr ? (d->ghost_ld != 0 ? static_cast<int64_t>(d->batch) * d->ghost_ld : 2 * d->g) : 0;
  }
  int C = g_reverse_variant;
  if (C == 0) {  // least redundant work; ties go to the shorter run (more warps per SM)
    C = 12;
    for (int c : {16, 20, 24})
      if (rev_work(d->n, c) < rev_work(d->n, C)) C = c;
  }
  switch (C) {
    case 12: return launch_reverse<12>(p, d->batch, st);
    case 16: return launch_reverse<16>(p, d->batch, st);
    case 20: return launch_reverse<20>(p, d->batch, st);
    default: return launch_reverse<24>(p, d->batch, st);
  }
}

}  // namespace psk

using namespace psk;

extern "C" {

int psk_set_reverse_variant(int variant) {
  if (variant != 0 && variant != 12 && variant != 16 && variant != 20 && variant != 24) return PSK_E_INVALID;
  g_reverse_variant = variant;
  return PSK_OK;
}

int psk_ssprk33_step_adjoint(const psk_desc *d, const double *u, const double *p_in, const double *dt,
                             int64_t dt_stride, double *p_out, double *k1_out, double *k2_out,
                             psk_stream_t stream) {
  return run_reverse(d, u, p_in, dt, dt_stride, nullptr, p_out, k1_out, k2_out, static_cast<cudaStream_t>(stream));
}

int psk_ssprk33_step_adjoint_bc(const psk_desc *d, const double *u, const double *p_in, const double *dt,
                                int64_t dt_stride, const double *ghost3, double *p_out, double *k1_out,
                                double *k2_out, psk_stream_t stream) {
  if (d != nullptr && d->bc != PSK_BC_DIRICHLET) return PSK_E_UNSUPPORTED;
  return run_reverse(d, u, p_in, dt, dt_stride, ghost3, p_out, k1_out, k2_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
