// psk_solve.cu -- the whole time loop of a small-grid solve in one launch (sm_100a, fp64).
#include <cuda_runtime.h>

#include <cstdint>

#include "psk_common.cuh"
#include "psk_math.cuh"

// ===========================================================================
// Whole time loop in one launch for rows that fit in shared memory (psk_solve_rows).
//
// The reference's own use cases are single small grids (examples/burgers.py: N = 256..512,
// 171 adaptive steps to t = 1): three launches plus a host round trip for dt per step make such
// runs launch-latency bound.  Here one CTA owns one row for the whole solve: u, k1, k2 and
// the face values live in shared memory, the CFL reduction (timestepping.py:139-142 with
// burgers/schemes.py:42-49, :121-127) is a block reduction, and nothing touches HBM between
// the initial load and the final store (except the optional per-step tape for the adjoint).
namespace psk {

struct SolveParams {
  double *u;          // [batch][ld] in / out
  double *t_out;      // [batch]
  int32_t *steps_out; // [batch]
  double *dt_hist;    // optional [batch][max_steps]
  double *tape;       // optional [(max_steps + 1)][batch][ld]: state before every step and the final one
  const double *nu;
  const double *vel;
  const double *vel_l;
  const double *vel_r;
  BcView bc;
  int64_t ld;
  int64_t tape_stride;  // batch * ld
  double dx, invdx, eps;
  double theta, cfl_scale, tfinal, fixed_dt;
  int adaptive, max_steps, batch;
  const double *dt_table;     // optional [max_steps]: the step sizes (fixed mode)
  const double *ghost_table;  // optional [max_steps][3][ghost_block]: boundary data at the three stage times
  int64_t ghost_block;        // doubles per (step, stage): 2 g (shared by all rows) or batch * ghost_ld
};

__device__ __forceinline__ double block_max_abs(const double *a, int lo, int hi, unsigned long long *scratch) {
  unsigned long long m = 0ull;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const unsigned long long b = abs_bits(a[i]);
    m = b > m ? b : m;
  }
  m = warp_max_bits(m);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) scratch[wid] = m;
  __syncthreads();
  unsigned long long r = 0ull;
  const int nw = (blockDim.x + 31) >> 5;
  for (int k = 0; k < nw; ++k) r = scratch[k] > r ? scratch[k] : r;
  return __longlong_as_double(static_cast<long long>(r));
}

template <int EQ, int FLUX, int REC, bool STRICT>
__global__ void __launch_bounds__(1024)
solve_rows_kernel(const SolveParams p) {
  extern __shared__ double smem[];
  const int row = blockIdx.x;
  const int g = p.bc.g, n = p.bc.n, nx = p.bc.nx;
  // five arrays of nx doubles: current state, two stage buffers, left / right face values
  // su / s1 swap roles after every step (stage 3 writes the new state over the dead k1)
  double *su = smem, *s1 = su + nx, *s2 = s1 + nx, *sl = s2 + nx, *sr = sl + nx;
  unsigned long long *scratch = reinterpret_cast<unsigned long long *>(sr + nx);
  double *grow = p.u + static_cast<int64_t>(row) * p.ld;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) su[i] = grow[i];
  __syncthreads();

  double t = 0.0;
  int m = 0;
  while (m < p.max_steps) {
    if (p.tape != nullptr) {
      double *trow = p.tape + static_cast<int64_t>(m) * p.tape_stride + static_cast<int64_t>(row) * p.ld;
      for (int i = threadIdx.x; i < nx; i += blockDim.x) trow[i] = su[i];
    }
    double dt;
    if (p.adaptive) {
      if (t >= p.tfinal) break;  // timestepping.py:133-134
      const double smax = block_max_abs(su, g, nx - g, scratch);
      dt = __dmul_rn(p.theta, __ddiv_rn(p.cfl_scale, smax));
      const double dt_min = __dadd_rn(p.tfinal, -t);
      dt = __dadd_rn(dt < dt_min ? dt : dt_min, 1.0e-15);  // timestepping.py:140-142
      if (!isfinite(dt)) {                                  // timestepping.py:144-145
        m = -1 - m;
        break;
      }
    } else {
      dt = (p.dt_table != nullptr) ? p.dt_table[m] : p.fixed_dt;
    }
    if (p.dt_hist != nullptr && threadIdx.x == 0) p.dt_hist[static_cast<int64_t>(row) * p.max_steps + m] = dt;

#pragma unroll 1
    for (int stage = 1; stage <= 3; ++stage) {
      const double *src = (stage == 1) ? su : (stage == 2 ? s1 : s2);
      double *dst = (stage == 2) ? s2 : s1;  // stage 1: k1 -> s1, stage 2: k2 -> s2, stage 3: u' -> s1
      // time-dependent boundary data: the values the user's g(t, x) takes at t, t + dt, t + dt / 2 (timestepping.py:314-319)
      BcView bc = p.bc;
      if (p.ghost_table != nullptr) bc.ghost = p.ghost_table + (static_cast<int64_t>(3) * m + (stage - 1)) * p.ghost_block;
      double speed = 0.0;
      if (FLUX == PSK_FLUX_LAX_FRIEDRICHS) {
        // max |w| over all cells after the boundary condition (scalar.py:277)
        unsigned long long mm = 0ull;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) {
          const unsigned long long b = abs_bits(load_w(bc, src, row, i));
          mm = b > mm ? b : mm;
        }
        mm = warp_max_bits(mm);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mm;
        __syncthreads();
        unsigned long long r = 0ull;
        for (int k = 0; k < static_cast<int>((blockDim.x + 31) >> 5); ++k) r = scratch[k] > r ? scratch[k] : r;
        speed = __longlong_as_double(static_cast<long long>(r));
      }
      // face values of every cell (zero padding beyond the array ends, BC in the ghost cells)
      for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const Weno5Pair o = reconstruct_cell<REC, STRICT>(
            load_w(bc, src, row, i - 2), load_w(bc, src, row, i - 1), load_w(bc, src, row, i),
            load_w(bc, src, row, i + 1), load_w(bc, src, row, i + 2), p.eps);
        sl[i] = o.ul;
        sr[i] = o.ur;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        double Flo = 0.0, Fhi = 0.0;  // jnp.pad(fnum, 1)
        if (i >= 1) {
          const int j = i - 1;
          const double nu = (p.nu != nullptr) ? p.nu[j] : 1.0;
          const double arj = (EQ != PSK_EQ_BURGERS) ? p.vel_r[j] : 0.0, alp = (EQ != PSK_EQ_BURGERS) ? p.vel_l[j + 1] : 0.0;
          Flo = face_flux<EQ, FLUX, STRICT>(sr[j], sl[j + 1], load_w(bc, src, row, j), load_w(bc, src, row, j + 1),
                                            speed, nu, arj, alp);
        }
        if (i <= nx - 2) {
          const int j = i;
          const double nu = (p.nu != nullptr) ? p.nu[j] : 1.0;
          const double arj = (EQ != PSK_EQ_BURGERS) ? p.vel_r[j] : 0.0, alp = (EQ != PSK_EQ_BURGERS) ? p.vel_l[j + 1] : 0.0;
          Fhi = face_flux<EQ, FLUX, STRICT>(sr[j], sl[j + 1], load_w(bc, src, row, j), load_w(bc, src, row, j + 1),
                                            speed, nu, arj, alp);
        }
        const double vel = (EQ == PSK_EQ_ADVECTION) ? p.vel[i] : 0.0;
        const double L = rhs_from_faces<EQ, STRICT>(Flo, Fhi, vel, p.dx, p.invdx);
        // like the reference, the combine uses the RAW stored value (not the boundary-filled one)
        dst[i] = stage_combine<STRICT>(stage, su[i], src[i], dt, L);
      }
      __syncthreads();
    }
    {
      double *tmp = su;  // the new state sits in s1
      su = s1;
      s1 = tmp;
    }
    t = __dadd_rn(t, dt);
    m += 1;
  }
  for (int i = threadIdx.x; i < nx; i += blockDim.x) grow[i] = su[i];
  if (p.tape != nullptr && m >= 0) {
    double *trow = p.tape + static_cast<int64_t>(m) * p.tape_stride + static_cast<int64_t>(row) * p.ld;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) trow[i] = su[i];
  }
  if (threadIdx.x == 0) {
    p.t_out[row] = t;
    p.steps_out[row] = m;
  }
}

template <int EQ, int FLUX, int REC, bool STRICT>
int launch_solve(const SolveParams &p, cudaStream_t st) {
  const int nx = p.bc.nx;
  const size_t smem = sizeof(double) * (5 * static_cast<size_t>(nx) + 40);
  if (smem > 227 * 1024) return PSK_E_UNSUPPORTED;
  int threads = ((nx + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 64) threads = 64;
  auto kern = solve_rows_kernel<EQ, FLUX, REC, STRICT>;
  PSK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<p.batch, threads, smem, st>>>(p);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

template <int EQ, int FLUX, bool STRICT>
int solve_rec(int rec, const SolveParams &p, cudaStream_t st) {
  switch (rec) {
    case PSK_REC_CONSTANT: return launch_solve<EQ, FLUX, PSK_REC_CONSTANT, STRICT>(p, st);
    case PSK_REC_WENOJS32: return launch_solve<EQ, FLUX, PSK_REC_WENOJS32, STRICT>(p, st);
    default: return launch_solve<EQ, FLUX, PSK_REC_WENOJS53, STRICT>(p, st);
  }
}

template <bool STRICT>
int solve_scheme(const psk_desc *d, const SolveParams &p, cudaStream_t st) {
  if (d->equation == PSK_EQ_ADVECTION) return solve_rec<PSK_EQ_ADVECTION, PSK_FLUX_UPWIND, STRICT>(d->rec, p, st);
  if (d->equation == PSK_EQ_CONTINUITY) return solve_rec<PSK_EQ_CONTINUITY, PSK_FLUX_UPWIND, STRICT>(d->rec, p, st);
  switch (d->flux) {
    case PSK_FLUX_RUSANOV: return solve_rec<PSK_EQ_BURGERS, PSK_FLUX_RUSANOV, STRICT>(d->rec, p, st);
    case PSK_FLUX_LAX_FRIEDRICHS: return solve_rec<PSK_EQ_BURGERS, PSK_FLUX_LAX_FRIEDRICHS, STRICT>(d->rec, p, st);
    case PSK_FLUX_UPWIND: return solve_rec<PSK_EQ_BURGERS, PSK_FLUX_UPWIND, STRICT>(d->rec, p, st);
    default: return solve_rec<PSK_EQ_BURGERS, PSK_FLUX_ENGQUIST_OSHER, STRICT>(d->rec, p, st);
  }
}

}  // namespace psk

extern "C" int psk_solve_rows(const psk_desc *d, double *u, int adaptive, double theta, double cfl_scale,
                              double tfinal, double fixed_dt, int max_steps, double *t_out,
                              int32_t *steps_out, double *dt_hist, double *tape, psk_stream_t stream) {
  int rc = psk::check_desc(d);
  if (rc != PSK_OK) return rc;
  if (u == nullptr || t_out == nullptr || steps_out == nullptr || max_steps <= 0) return PSK_E_INVALID;
  if (adaptive && d->equation != PSK_EQ_BURGERS) return PSK_E_UNSUPPORTED;  // state-independent dt: pass it fixed
  // no single-launch form of ESWENO32: callers fall back to the step-by-step path
  if (d->rec == PSK_REC_ESWENO32 || d->flux == PSK_FLUX_ESWENO) return PSK_E_UNSUPPORTED;
  psk::SolveParams p{};
  p.u = u; p.t_out = t_out; p.steps_out = steps_out; p.dt_hist = dt_hist; p.tape = tape;
  p.nu = d->nu; p.vel = d->velocity; p.vel_l = d->vel_l; p.vel_r = d->vel_r;
  p.bc = psk::make_bc_view(d);
  p.ld = d->ld;
  p.tape_stride = static_cast<int64_t>(d->batch) * d->ld;
  p.dx = d->dx; p.invdx = 1.0 / d->dx; p.eps = d->eps;
  p.theta = theta; p.cfl_scale = cfl_scale; p.tfinal = tfinal; p.fixed_dt = fixed_dt;
  p.adaptive = adaptive; p.max_steps = max_steps; p.batch = d->batch;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d->math == PSK_MATH_STRICT ? psk::solve_scheme<true>(d, p, st) : psk::solve_scheme<false>(d, p, st);
}

/* psk_solve_rows for step sizes and boundary data that are known in advance but change from step to step
 * (state-independent dt clamped at tfinal, time-dependent Dirichlet data: drivers/advection-adjoint.py): */
extern "C" int psk_solve_rows_tables(const psk_desc *d, double *u, int nsteps, const double *dt_table,
                                     const double *ghost_table, double *t_out, int32_t *steps_out, double *tape,
                                     psk_stream_t stream) {
  int rc = psk::check_desc(d);
  if (rc != PSK_OK) return rc;
  if (u == nullptr || t_out == nullptr || steps_out == nullptr || nsteps <= 0 || dt_table == nullptr) return PSK_E_INVALID;
  if (d->rec == PSK_REC_ESWENO32 || d->flux == PSK_FLUX_ESWENO) return PSK_E_UNSUPPORTED;
  psk::SolveParams p{};
  p.u = u; p.t_out = t_out; p.steps_out = steps_out; p.tape = tape;
  p.nu = d->nu; p.vel = d->velocity; p.vel_l = d->vel_l; p.vel_r = d->vel_r;
  p.bc = psk::make_bc_view(d);
  p.ld = d->ld;
  p.tape_stride = static_cast<int64_t>(d->batch) * d->ld;
  p.dx = d->dx; p.invdx = 1.0 / d->dx; p.eps = d->eps;
  p.max_steps = nsteps; p.batch = d->batch;
  p.dt_table = dt_table;
  p.ghost_table = ghost_table;
  p.ghost_block = d->ghost_ld != 0 ? static_cast<int64_t>(d->batch) * d->ghost_ld : 2 * d->g;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d->math == PSK_MATH_STRICT ? psk::solve_scheme<true>(d, p, st) : psk::solve_scheme<false>(d, p, st);
}

// ===========================================================================
// FP64 peak probe for the roofline: 16 independent DFMA chains per thread, enough CTAs to fill
// every SM.  MEASURED_PEAKS.json has no fp64 figure, bench.py measures it with this kernel.
namespace psk {
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = threadIdx.x * 1e-3 + k;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) acc += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
}  // namespace psk

/* Launch the DFMA probe: ctas * 256 threads, each doing iters * 64 DFMAs (16 independent chains); out needs ctas * 256
 * doubles.  The caller times it (CUDA events) and computes 2 * 64 * iters * threads / seconds. */
extern "C" int psk_dfma_probe(double *out, int ctas, int iters, psk_stream_t stream) {
  if (out == nullptr || ctas <= 0 || iters <= 0) return PSK_E_INVALID;
  psk::dfma_probe_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, iters, 0.999999, 1.0e-6);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}
