/*
 * psk.h -- C ABI of libpsk, the B200-native (sm_100a, fp64) implementation of
 * the one data-parallel hot path of alexfikl/pyshocks:
 *
 *   apply_boundary -> WENO-JS reconstruction -> numerical flux -> flux-difference
 *   RHS, fused with the SSPRK33 stage update; the CFL max-wavespeed reduction;
 *   and the hand-derived discrete adjoint of the stage.
 *
 * The reference has no FFI: its extension point is functools.singledispatch
 * registration on scheme / stepper / boundary types.
 * Each entry point below names the reference function it stands in for
 * (paths relative to /root/reference/src/pyshocks).  The Python shims in
 * pyshocks_b200/ register these behind the same generic functions and call
 * them through ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - every array is fp64 in DEVICE memory, owned by the caller;
 *  - a "state" array holds `batch` rows of `nx = n + 2 g` cells, row stride
 *    `ld` doubles (ld >= nx); row r, cell i lives at base[r * ld + i]; cells
 *    [0, g) and [nx - g, nx) are ghost cells (grid.py:83-117);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    allocates nothing and never synchronises unless its name ends in _sync;
 *  - return value: PSK_OK or a psk_status code (never a CPU fallback).
 */
#ifndef PSK_H
#define PSK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSK_VERSION 104 /* 0.1.4: psk_ssprk33_step_adjoint_bc */

typedef void *psk_stream_t; /* cudaStream_t */

enum psk_status {
  PSK_OK = 0,
  PSK_E_INVALID = 1,     /* bad argument (NULL pointer, n <= 0, g too small, ...)   */
  PSK_E_UNSUPPORTED = 2, /* valid in pyshocks but outside the hot-path scope         */
  PSK_E_CUDA = 3,        /* a CUDA runtime call failed; see psk_last_cuda_error()    */
  PSK_E_NONFINITE = 4    /* non-finite time step (timestepping.py:144-145)           */
};

/* equation family: burgers/schemes.py, advection/schemes.py, continuity/schemes.py */
enum psk_equation { PSK_EQ_BURGERS = 0, PSK_EQ_ADVECTION = 1, PSK_EQ_CONTINUITY = 2 };

/* numerical flux (scalar.py:91-322; names of burgers/__init__.py:62-72)
 *   RUSANOV          scalar_flux_rusanov        (local Lax-Friedrichs, "rusanov")
 *   LAX_FRIEDRICHS   scalar_flux_lax_friedrichs (global max speed,     "lf")
 *   UPWIND           scalar_flux_upwind / upwind_flux (averaged-speed switch, "godunov")
 *   ENGQUIST_OSHER   scalar_flux_engquist_osher (omega = 0,            "eo")      */
enum psk_flux {
  PSK_FLUX_RUSANOV = 0,
  PSK_FLUX_LAX_FRIEDRICHS = 1,
  PSK_FLUX_UPWIND = 2,
  PSK_FLUX_ENGQUIST_OSHER = 3,
  /* Burgers ESWENO32 scheme (burgers/schemes.py:205-256, "esweno32"): the UPWIND flux plus the
   * dissipative flux built from the ESWENO weights; needs PSK_REC_ESWENO32 */
  PSK_FLUX_ESWENO = 4
};

/* reconstruction.py:127-163 (constant), :319-348 (WENOJS32 / WENOJS53), :386-439 (ESWENO32:
 * the JS-3 stencils with the weights of weno.py:284-296; desc.eps and desc.delta carry its
 * parameters, weno.py:263-281) */
enum psk_rec { PSK_REC_CONSTANT = 0, PSK_REC_WENOJS32 = 1, PSK_REC_WENOJS53 = 2, PSK_REC_ESWENO32 = 3 };

/* ghost-cell fill applied at the top of every RHS (schemes.py:343)
 *   PERIODIC  scalar.py:529-540
 *   DIRICHLET scalar.py:418-427; desc.ghost holds the 2 g values g(t, x_ghost),
 *             left ghosts first, evaluated on the host by the caller
 *   NEUMANN   scalar.py:472-500; ghost = mirrored interior cell + desc.ghost[k],
 *             desc.ghost holding side * (x[ifrom] - x[ito]) * g(t) per ghost cell
 *   NONE      ghost cells are used as found (halo-exchanged slabs)              */
enum psk_bc { PSK_BC_PERIODIC = 0, PSK_BC_DIRICHLET = 1, PSK_BC_NEUMANN = 2, PSK_BC_NONE = 3 };

/* arithmetic contract
 *   STRICT  every operation in the reference's order, no FMA contraction, IEEE
 *           division: results are bit-identical to oracle/psk_oracle.c
 *   FAST    same algorithm, re-associated for the FP64 pipe (FMA, shared
 *           smoothness indicators, one division per reconstructed value);
 *           agrees with STRICT to a few ulp per RHS evaluation               */
enum psk_math { PSK_MATH_FAST = 0, PSK_MATH_STRICT = 1 };

typedef struct psk_desc {
  int32_t equation; /* enum psk_equation */
  int32_t flux;     /* enum psk_flux     */
  int32_t rec;      /* enum psk_rec      */
  int32_t bc;       /* enum psk_bc       */
  int32_t math;     /* enum psk_math     */
  int32_t n;        /* interior cells per row                                    */
  int32_t g;        /* ghost cells per side, >= stencil width of rec (1 / 2 / 3) */
  int32_t batch;    /* rows (independent problems; 1 for a plain pyshocks array) */
  int64_t ld;       /* row stride of state arrays, in doubles                    */
  double dx;        /* cell size h; grid.dx is a constant array (grid.py:164)    */
  double eps;       /* WENO-JS epsilon (reconstruction.py:325, :341)             */
  const double *nu; /* NULL: nu = 1 (|alpha - 1| <= 1e-8); else nx - 1 per-face
                       values grid.df ** (alpha - 1) (scalar.py:231-234)         */
  const double *velocity; /* nx, advection / continuity only, shared by all rows */
  const double *vel_l;    /* nx, left  values of reconstruct(velocity)  ("al")   */
  const double *vel_r;    /* nx, right values of reconstruct(velocity)  ("ar")   */
  const double *ghost;    /* DIRICHLET / NEUMANN data, 2 g doubles per row       */
  int64_t ghost_ld;       /* row stride of ghost in doubles; 0 = shared          */
  double delta;           /* ESWENO32 delta (reconstruction.py:396; burgers/schemes.py:241) */
} psk_desc;

int psk_version(void);
const char *psk_status_string(int status);
/* last cudaError_t seen by this library on the calling thread (0 = none) */
int psk_last_cuda_error(void);
/* A/B switch between the two implementations of the fused stage (same results):
 * 0 = warp-shuffle kernel (default), 1 = shared-memory tile kernel. */
int psk_set_stage_variant(int variant);
/* Same for the adjoint stage: 0 = warp-shuffle kernel where applicable (default), 1 = tile kernel. */
int psk_set_adjoint_variant(int variant);

/* apply_boundary(bc, grid, t, u) -> w            schemes.py:431-442, scalar.py BCs.
 * w may alias u. */
int psk_apply_boundary(const psk_desc *d, const double *u, double *w, psk_stream_t stream);

/* reconstruct(rec, grid, bc, f, f, .) -> (fl, fr)   reconstruction.py:79-118, :358-377.
 * Zero-padded beyond the array ends like jnp.convolve(..., "same") (convolve.py:113-114). */
int psk_reconstruct(const psk_desc *d, const double *f, double *fl, double *fr,
                    psk_stream_t stream);

/* numerical_flux(scheme, grid, bc, t, w) -> F[nx + 1]       schemes.py:321-336 and the
 * registrations burgers/schemes.py:83-196, advection/schemes.py:121-129,
 * continuity/schemes.py:93-110.  w already has its ghost cells set; F has row
 * stride ld_f >= nx + 1 and F[0] = F[nx] = 0 (the reference's jnp.pad).
 * lf_work: [batch] doubles of scratch, required for PSK_FLUX_LAX_FRIEDRICHS (receives the
 * global speed max |w|, scalar.py:277), may be NULL otherwise. */
int psk_numerical_flux(const psk_desc *d, const double *w, double *flux, int64_t ld_f,
                       double *lf_work, psk_stream_t stream);

/* apply_operator(scheme, grid, bc, t, u) -> L[nx]     schemes.py:339-346,
 * advection/schemes.py:62-73.  Applies the boundary condition internally and
 * evaluates ALL nx rows, ghost rows included, exactly like the reference.
 * rhs must not alias u.  lf_work: see psk_numerical_flux. */
int psk_apply_operator(const psk_desc *d, const double *u, double *rhs, double *lf_work,
                       psk_stream_t stream);

/* max |u| per row -> out[batch].  mode 1: over the interior (the reduction inside
 * predict_timestep, burgers/schemes.py:47, :124); mode 0: over all nx stored cells;
 * mode 2: over all nx cells after the boundary condition of d (the global
 * Lax-Friedrichs speed, scalar.py:277).  NaN propagates like jnp.max. */
int psk_max_abs(const psk_desc *d, const double *u, int mode, double *out,
                psk_stream_t stream);

/* One fused SSPRK33 stage (timestepping.py:312-320 with apply_operator inlined):
 *   stage 1: uout = uin + dt L(uin)                     (uin = u0)
 *   stage 2: uout = 3/4 u0 + 1/4 (uin + dt L(uin))
 *   stage 3: uout = 1/3 u0 + 2/3 (uin + dt L(uin))
 *   stage 0: uout = L(uin)                              (plain RHS, interior fast path)
 * dt: device pointer, row r uses dt[r * dt_stride] (dt_stride 0 = one shared value).
 * active: optional per-row byte mask; rows with active[r] == 0 are left untouched.
 * maxabs: optional [batch]; on return max over the interior of |uout| per row
 *         (fused CFL reduction for the next predict_timestep); needs no pre-zeroing
 *         ordering beyond being zero-filled by the caller before the launch.
 * lf_work: see psk_numerical_flux (one extra reduction pass over uin).
 * uout must not alias uin (it may alias u0 in stage 3).
 * ghost_rows != 0 also produces the nx - n ghost entries of uout exactly as the
 * reference does (zero-padded stencils at the array ends); otherwise ghost cells
 * of uout are not written. */
int psk_ssprk33_stage(const psk_desc *d, int stage, const double *u0, const double *uin,
                      double *uout, const double *dt, int64_t dt_stride,
                      const uint8_t *active, double *lf_work, double *maxabs,
                      int ghost_rows, psk_stream_t stream);

/* The fused right-hand side with a GENERAL combine,
 *     uout = ca u0 + cb uin + cc dt L(uin),        L = apply_operator (schemes.py:339-346),
 * from which the stages of the other steppers of the reference are built (timestepping.py:289-405): ForwardEuler
 * (1, 0, 1 with u0 = uin), RK44 (y2 = u + dt/2 L(u), ..., u' = acc + 1/3 y4 + dt/6 L(y4)), CKRK45
 * (k = a_i k + dt L(p)).  Every scheme, reconstruction, boundary kind and math mode of psk_ssprk33_stage;
 * uout must not alias uin (it may alias u0).  lf_work, ghost_rows as in psk_ssprk33_stage. */
int psk_rhs_axpby(const psk_desc *d, const double *u0, const double *uin, double *uout, const double *dt,
                  int64_t dt_stride, double ca, double cb, double cc, double *lf_work, int ghost_rows,
                  psk_stream_t stream);

/* psk_ssprk33_stage for the global Lax-Friedrichs flux (scalar.py:258-278) on PERIODIC rows WITHOUT the reduction
 * pass: there the speed max |w| over all cells after the boundary condition (scalar.py:277) is max |uin| over the
 * interior (the ghost cells are copies), which the launch that PRODUCED uin has already reduced into its `maxabs`
 * output.  speed [batch]: that value (for the first stage of a solve: psk_max_abs, interior); maxabs [batch],
 * zero-filled by the caller: max |uout| over the interior = the speed of the next stage.  Three launches per step
 * instead of six, one full read of the state less per stage; same bits.  PSK_E_UNSUPPORTED for other fluxes and
 * boundary kinds. */
int psk_ssprk33_stage_lf(const psk_desc *d, int stage, const double *u0, const double *uin, double *uout,
                         const double *dt, int64_t dt_stride, const double *speed, double *maxabs,
                         psk_stream_t stream);

/* One whole SSPRK33 step (timestepping.py:312-320: the three stages above) in ONE launch, for the
 * hot configuration only: Burgers with the Rusanov (nu = 1), upwind or Engquist-Osher flux + WENO-JS5,
 * FAST math, periodic rows (g >= 3) or slabs of a larger grid (boundary kind NONE with g >= 9 ghost
 * cells per side, filled by the caller before the call: the three stages reach 9 cells beyond the row),
 * 16-byte aligned rows.  Also the global Lax-Friedrichs flux (scalar.py:258-278, nu = 1) on PERIODIC rows of at
 * most 16 512 cells: one thread-block cluster per row, the row-wide max |w| of every stage input (scalar.py:277)
 * exchanged through distributed shared memory.  u is read once and uout written once; the stage values stay in registers
 * (temporal blocking, psk_fast_kernels.cuh).  Bit-identical to three psk_ssprk33_stage calls.
 * uout must not alias u.  active / maxabs as in psk_ssprk33_stage, except that rows with
 * active[r] == 0 are COPIED to uout (the state ping-pongs between two arrays).  Ghost cells of uout
 * are not written.  PSK_E_UNSUPPORTED outside that configuration: call psk_ssprk33_stage three
 * times instead.
 * Advection / continuity on PERIODIC rows (upwind flux) are covered too, under one condition the caller
 * guarantees: the velocity's reconstruction is periodic like the state, vel_r[g - 1] == vel_r[g + n - 1] and
 * vel_l[g + n] == vel_l[g] (true whenever the ghost cells of the velocity array are the periodic images of its
 * interior).  A window cell beyond a row end is advanced with the velocity data of the interior cell it is the image
 * of; with that condition those are the numbers the stage kernels read at the row ends, hence the same bits. */
int psk_ssprk33_step(const psk_desc *d, const double *u, double *uout, const double *dt,
                     int64_t dt_stride, const uint8_t *active, double *maxabs, psk_stream_t stream);

/* psk_ssprk33_step for rows with DIRICHLET or NEUMANN boundary data (scalar.py:418-427, 472-500) and for the
 * advection / continuity equations: the whole step in one launch, the ghost cells of every stage input taking the
 * data of that stage's time (Neumann: plus the stage value of their mirror image).  ghost3: three consecutive blocks of (d->ghost_ld != 0 ? batch * d->ghost_ld : 2 g) doubles -- what the
 * user's g(t, x) returns at t, t + dt, t + dt / 2 (timestepping.py:314-319), rows at stride d->ghost_ld (0: one
 * set for all rows), left ghost cells first; d->ghost is ignored.  Burgers + {Rusanov, upwind,
 * Engquist-Osher; global Lax-Friedrichs on Dirichlet rows of at most 16 512 cells},
 * advection, continuity (upwind flux); WENO-JS5, FAST math, 16-byte aligned rows, g <= 16.  d->nu (Rusanov /
 * Lax-Friedrichs with alpha != 1, scalar.py:231-234): the viscosity of every face of the array, nx - 1 values.
 * PSK_E_UNSUPPORTED elsewhere.  Same bits as three psk_ssprk33_stage calls with those data.  active / maxabs
 * as in psk_ssprk33_step.  k1_out / k2_out (both or neither; then active = maxabs = NULL): the stage values are
 * stored as well, as psk_ssprk33_step_stages does for periodic rows, and uout may be NULL (third stage skipped):
 * the recomputation of the reverse sweep in one launch. */
int psk_ssprk33_step_bc(const psk_desc *d, const double *u, double *uout, const double *dt, int64_t dt_stride,
                        const double *ghost3, const uint8_t *active, double *maxabs, double *k1_out, double *k2_out,
                        psk_stream_t stream);

/* `nsteps` whole-step launches (psk_ssprk33_step, or psk_ssprk33_step_bc for Dirichlet rows) enqueued back to
 * back from one call, every state written straight onto the tape: tape[0 * tape_stride ..] holds the initial state,
 * on return tape[m * tape_stride ..] the state after m steps -- the InMemoryCheckpoint contents of
 * timestepping.step (timestepping.py:130-131) for step sizes known in advance, without a copy.  dt_table,
 * ghost_table as in psk_solve_rows_tables.  PSK_E_UNSUPPORTED (nothing launched) where the whole-step kernel
 * does not exist. */
int psk_ssprk33_steps_tape(const psk_desc *d, double *tape, int64_t tape_stride, int nsteps, const double *dt_table,
                           const double *ghost_table, psk_stream_t stream);

/* psk_ssprk33_step that also stores the stage values k1, k2 (timestepping.py:314-317) -- what the reverse
 * sweep recomputes from a checkpointed state before its three psk_ssprk33_stage_adjoint calls -- in one
 * launch instead of two or three.  uout may be NULL (only k1, k2 wanted: the third stage is skipped).
 * Burgers + Rusanov + WENO-JS5, FAST math, periodic rows, aligned rows; PSK_E_UNSUPPORTED elsewhere.
 * Same bits as psk_ssprk33_stage.  No array may alias another. */
int psk_ssprk33_step_stages(const psk_desc *d, const double *u, double *k1, double *k2, double *uout,
                            const double *dt, int64_t dt_stride, psk_stream_t stream);

/* One whole REVERSE SSPRK33 step in ONE launch (psk_reverse_kernels.cuh): what adjoint_step computes per
 * step as jax.jacfwd(advance)(dt, t, u).T @ p (timestepping.py:174, :198-209),
 *     p_out = (d advance(dt, u) / d u)^T p_in,
 * from the checkpointed state u alone: the stage values k1, k2 are recomputed inside the kernel, then the
 * three adjoint stages of psk_ssprk33_stage_adjoint follow, with k1, k2, lam2, lam1 in shared memory
 * (24 B of DRAM traffic per cell instead of the 144 B of five launches).  Rows are periodic RINGS of
 * the n interior cells: only interior cells of u, p_in are read and only interior cells of p_out are
 * written (the gradient of the interior -> interior map; no cotangent lives on ghost cells).
 * Boundary kind NONE with g >= 16 ghost cells: the row is a SLAB of a larger grid, the ghost cells of u and p_in
 * hold the neighbouring slabs' edge cells (the caller's exchange, e.g. psk_halo_push) and p_out is the slab's
 * part of the gradient -- the transposed stencil needs no "send back and add", every slab gathers.
 * Burgers + Rusanov (nu = 1) + WENO-JS5, FAST math, periodic or NONE (DIRICHLET: see below), 16-byte aligned rows, n even;
 * PSK_E_UNSUPPORTED elsewhere (recompute with psk_ssprk33_step_stages and call
 * psk_ssprk33_stage_adjoint three times instead).  k1_out / k2_out: optional (both or neither), the
 * recomputed stage values of the interior cells, bit-identical to psk_ssprk33_stage.
 * p_out must not alias p_in or u. */
int psk_ssprk33_step_adjoint(const psk_desc *d, const double *u, const double *p_in, const double *dt,
                             int64_t dt_stride, double *p_out, double *k1_out, double *k2_out,
                             psk_stream_t stream);

/* psk_ssprk33_step_adjoint on DIRICHLET rows (scalar.py:418-427; the boundary kind of the reference's own
 * burgers-adjoint template, drivers/burgers-adjoint.py:68-97): the ghost cells of the stage inputs u, k1, k2 take
 * the data of the stage times t, t + dt, t + dt / 2 (ghost3: three blocks laid out as in psk_ssprk33_step_bc;
 * NULL: d->ghost for all three), and since these data do not depend on the state no cotangent lives on them:
 *     p_out = (d advance(dt, u)[interior] / d u[interior])^T p_in[interior].
 * Only interior cells of p_in are read (its ghost cells count as zero -- what apply_boundary with homogeneous
 * Dirichlet data makes of the adjoint variable after every step, timestepping.py:208-209) and only interior
 * cells of p_out are written.  g = 3; otherwise the conditions, options and status codes of
 * psk_ssprk33_step_adjoint, whose plain form also accepts DIRICHLET rows with d->ghost. */
int psk_ssprk33_step_adjoint_bc(const psk_desc *d, const double *u, const double *p_in, const double *dt,
                                int64_t dt_stride, const double *ghost3, double *p_out, double *k1_out,
                                double *k2_out, psk_stream_t stream);

/* A/B switch of psk_ssprk33_step_adjoint: cells per lane of its windows (12, 16, 20, 24; 0 = automatic). */
int psk_set_reverse_variant(int variant);

/* Device-side step control of timestepping.step (timestepping.py:128-150) for
 * Burgers-type schemes, per row r:
 *   dt   = theta * (cfl_scale / maxabs[r])             (burgers/schemes.py:42-49, :121-127;
 *                                                        cfl_scale = 0.5 dx_min^(2 - alpha))
 *   dt   = min(dt, tfinal - t[r]) + 1e-15 ; t[r] += dt on the NEXT call (see t_next)
 *   active[r] = t[r] < tfinal ;  *nonfinite |= !isfinite(dt)
 * Writes dt[r]; t_next[r] = t[r] + dt for active rows. */
int psk_step_control(int32_t batch, double theta, double cfl_scale, double tfinal,
                     const double *maxabs,
                     const double *t, double *t_next, double *dt, uint8_t *active,
                     int32_t *nonfinite, psk_stream_t stream);

/* The whole loop of timestepping.step (timestepping.py:128-152) with advance(SSPRK33) in ONE
 * launch, one CTA per row, for rows that fit in shared memory (5 nx doubles <= 227 KB): the
 * reference's own small-grid runs (examples/burgers.py) are launch-latency bound otherwise.
 *   adaptive != 0: Burgers schemes, dt = theta * (cfl_scale / max|u_interior|), clamped to tfinal,
 *                  + 1e-15, until t >= tfinal or max_steps;   adaptive == 0: max_steps steps of fixed_dt.
 * Boundary data (d->ghost) must not depend on time.  u is advanced in place (all nx entries,
 * ghost rows exactly like the reference).  t_out[batch], steps_out[batch] (negative: non-finite
 * dt at step -1 - value); dt_hist: optional [batch][max_steps]; tape: optional
 * [(max_steps + 1)][batch][ld] receiving the state before every step and the final state
 * (the InMemoryCheckpoint contents, timestepping.py:130-131).  PSK_E_UNSUPPORTED if a row does
 * not fit in shared memory. */
int psk_solve_rows(const psk_desc *d, double *u, int adaptive, double theta, double cfl_scale,
                   double tfinal, double fixed_dt, int max_steps, double *t_out,
                   int32_t *steps_out, double *dt_hist, double *tape, psk_stream_t stream);

/* psk_solve_rows for step sizes and boundary data that are known in advance but change from step to step:
 * `nsteps` steps with dt_table[m] (device, [nsteps]) -- e.g. the state-independent time step of the advection
 * and continuity schemes clamped at tfinal (timestepping.py:139-142 with advection/schemes.py:51-59) -- and, if
 * ghost_table != NULL, the Dirichlet / Neumann data of step m at the stage times t, t + dt, t + dt / 2
 * (timestepping.py:314-319: what the user's g(t, x) returns, evaluated by the caller) as
 * ghost_table[(3 m + stage) * block + ...], block = 2 g (d->ghost_ld == 0: shared by all rows) or
 * batch * d->ghost_ld.  drivers/advection-adjoint.py (BASELINE config 2) in one launch.  tape as above. */
int psk_solve_rows_tables(const psk_desc *d, double *u, int nsteps, const double *dt_table,
                          const double *ghost_table, double *t_out, int32_t *steps_out, double *tape,
                          psk_stream_t stream);

/* The whole reverse sweep of adjoint_step (timestepping.py:155-215) in ONE call, from the tape of
 * psk_solve_rows / psk_solve_rows_tables (state m at tape + m * tape_stride): for m = nsteps - 1 .. 0
 *     p <- (d advance(dt_m, t_m, u_m) / d u)^T p,   then   p <- apply_boundary(pbc, p)   (pbc != NULL)
 * with the kernels of psk_ssprk33_stage (recomputation of k1, k2, ghost rows included) and
 * psk_ssprk33_stage_adjoint, enqueued back to back (6 launches per step, no host round trip).
 * dt_table, ghost_table as in psk_solve_rows_tables.  pbc: descriptor of the boundary condition imposed on the
 * adjoint variable after every step (the apply_boundary argument of adjoint_step; its data must not depend on
 * time), same n, g, batch, ld as d.  p [batch][ld]: in p(T) (already passed through pbc), out p(0).
 * states: scratch, 5 state arrays; work / lf_work as in psk_ssprk33_stage_adjoint / psk_ssprk33_stage.
 * p_hist: optional [nsteps][batch][ld], p after step m (every AdjointStepCompleted.p). */
int psk_ssprk33_adjoint_sweep(const psk_desc *d, const double *tape, int64_t tape_stride, int nsteps,
                              const double *dt_table, const double *ghost_table, const psk_desc *pbc, double *p,
                              double *states, double *work, double *lf_work, double *p_hist, psk_stream_t stream);

/* FP64 peak probe for the roofline (not part of the reference-facing surface): ctas x 256 threads,
 * each executing iters x 64 DFMAs in 16 independent chains; out: ctas x 256 doubles. */
int psk_dfma_probe(double *out, int ctas, int iters, psk_stream_t stream);

/* out = J_L(u)^T v, the vector-Jacobian product of apply_operator w.r.t. u (all nx
 * rows, boundary condition included); what jax.vjp(apply_operator) returns, and the
 * building block of the reference's adjoint_step (timestepping.py:174, :205-206).  Every equation, flux and
 * reconstruction of the forward entry points, the ESWENO32 reconstruction and the Burgers ESWENO32 scheme included.
 * work: batch * (2 g + 3) doubles of scratch.  out must not alias u or v. */
int psk_apply_operator_vjp(const psk_desc *d, const double *u, const double *v, double *out,
                           double *work, psk_stream_t stream);

/* One fused adjoint stage:  out = c_acc * acc + c_acc2 * acc2 + c_v * (v + dt J_L(x)^T v)
 * The three calls of a reverse SSPRK33 step, given the recomputed
 * forward stages k1, k2 of the checkpointed state u and the incoming adjoint p':
 *   lam2 = 2/3 (p' + dt J(k2)^T p')
 *   lam1 = 1/4 (lam2 + dt J(k1)^T lam2)
 *   p    = 1/3 p' + 3/4 lam2 + (lam1 + dt J(u)^T lam1)
 * acc / acc2 may be NULL (their coefficient is then ignored).  dt as in
 * psk_ssprk33_stage.  work: batch * (2 g + 3) doubles of scratch.  out must not alias
 * x or v. */
int psk_ssprk33_stage_adjoint(const psk_desc *d, const double *x, const double *v,
                              const double *dt, int64_t dt_stride, double c_v,
                              const double *acc, double c_acc, const double *acc2,
                              double c_acc2, double *work, double *out,
                              psk_stream_t stream);

/* ---- slab decomposition of one large grid over several GPUs (BASELINE.json configs[3]) ----
 * The reference is single-device (pyshocks/__init__.py:66); the only reference notion involved is
 * the ghost-cell layout of grid.py:83-117: a slab array is [g | n_local | g] and its ghost cells
 * are the neighbours' edge cells.  One process per GPU; the neighbours' arrays are mapped
 * through CUDA IPC and written directly over NVLink. */
#define PSK_IPC_HANDLE_BYTES 64

/* cudaMalloc + zero-fill of `bytes` of peer-visible device memory on the current device;
 * `handle` receives PSK_IPC_HANDLE_BYTES bytes to pass to the other processes.  Synchronous. */
int psk_p2p_alloc(uint64_t bytes, void **ptr, unsigned char *handle);
int psk_p2p_free(void *ptr);
/* Map another process' psk_p2p_alloc allocation into this process (peer access is enabled
 * lazily); *ptr is its base address here.  Synchronous. */
int psk_p2p_open(const unsigned char *handle, void **ptr);
int psk_p2p_close(void *ptr);

/* Ghost-cell push: copy `count` doubles src_lo -> dst_lo (my first interior cells into the LEFT
 * neighbour's right ghost slots) and src_hi -> dst_hi (my last interior cells into the RIGHT
 * neighbour's left ghost slots), order the stores system-wide, then store `epoch` into
 * *flag_lo and *flag_hi (flags living in the neighbours' memory).  Any of the two sides may be
 * NULL.  One tiny launch. */
int psk_halo_push(const double *src_lo, double *dst_lo, const double *src_hi, double *dst_hi,
                  int32_t count, int64_t *flag_lo, int64_t *flag_hi, int64_t epoch,
                  psk_stream_t stream);

/* Stream-ordered wait: the launch completes once *flag_a >= epoch and *flag_b >= epoch (local
 * flags written by the neighbours' psk_halo_push; either may be NULL), or after timeout_ns,
 * in which case *timed_out is set to 1 (never a hang). */
int psk_halo_wait(const int64_t *flag_a, const int64_t *flag_b, int64_t epoch, int64_t timeout_ns,
                  int32_t *timed_out, psk_stream_t stream);

/* The same exchange FUSED into the stage kernel (one launch per stage, nothing else on the
 * exchange path): the warps that read ghost cells spin on the local flags wait_lo / wait_hi
 * until they reach wait_epoch (the neighbours' pushes of uin's edge cells), every other warp
 * starts at once; the lanes that store uout's first / last g cells also store them to
 * peer_lo / peer_hi (the neighbours' ghost slots of their uout array) and then raise
 * *flag_lo / *flag_hi to wait_epoch + 1.  NULL pointers switch a side off. */
typedef struct psk_halo_link {
  const int64_t *wait_lo, *wait_hi; /* LOCAL flags, written by the left / right neighbour        */
  int64_t wait_epoch;
  double *peer_lo, *peer_hi;        /* left neighbour's right ghost slots / right neighbour's left */
  int64_t *flag_lo, *flag_hi;       /* the neighbours' flags this rank raises                     */
  int64_t timeout_ns;               /* a spin gives up after this long and sets *timed_out        */
  int32_t *timed_out;
  /* Optional device-side epoch (for time loops replayed as CUDA graphs, where launch arguments
   * cannot change): when epoch_in != NULL the epoch is *epoch_in + wait_epoch, and one lane
   * stores epoch + 1 to *epoch_out for the next stage.  epoch_in / epoch_out must be two
   * different slots (stage k reads slot k & 1 and writes slot (k + 1) & 1). */
  const int64_t *epoch_in;
  int64_t *epoch_out;
} psk_halo_link;

/* psk_ssprk33_stage (stages 1-3, one row, boundary kind NONE, dt shared) with the exchange
 * above.  PSK_E_UNSUPPORTED outside the hot configuration (Burgers + Rusanov + WENO-JS5, FAST
 * math, g = 3, n % 4 == 0, aligned rows): fall back to psk_halo_wait / psk_ssprk33_stage /
 * psk_halo_push. */
int psk_ssprk33_stage_p2p(const psk_desc *d, int stage, const double *u0, const double *uin,
                          double *uout, const double *dt, double *maxabs, const psk_halo_link *link,
                          psk_stream_t stream);

/* psk_ssprk33_step (the whole SSPRK33 step in one launch) FUSED with the ghost-cell exchange of a
 * slab-decomposed grid: one row, boundary kind NONE, g = 9 ghost cells per side (the three stages of a step
 * reach 9 cells beyond the slab).  The warps whose windows reach into ghost cells wait on the local flags
 * link->wait_lo / wait_hi (>= wait_epoch) inside the kernel; the lanes that store the slab's first / last 9
 * cells of uout also store them to link->peer_lo / peer_hi (the neighbours' ghost slots of the array the
 * new state lives in, mapped with psk_p2p_open) and raise link->flag_lo / flag_hi to wait_epoch + 1.
 * epoch_in / epoch_out must be NULL.  Same bits as psk_ssprk33_step.  PSK_E_UNSUPPORTED outside Burgers +
 * {Rusanov (nu = 1), upwind, Engquist-Osher} + WENO-JS5 + FAST math, or for slabs shorter than 344 cells
 * (callers fall back to psk_halo_wait -> psk_ssprk33_step -> psk_halo_push). */
int psk_ssprk33_step_p2p(const psk_desc *d, const double *u, double *uout, const double *dt, double *maxabs,
                         const psk_halo_link *link, psk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PSK_H */
