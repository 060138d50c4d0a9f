/*
 * psk_oracle.c -- plain-C CPU restatement of the pyshocks hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/libpsk_oracle.so
 * and loaded only by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py.  It is never linked into libpsk.so.
 *
 * Pinning: checked bit-for-bit against oracle/pyshocks_oracle.py (NumPy), which in
 * turn is bit-for-bit equal to the golden vectors recorded from the reference's
 * own Python (tests/golden, .npz files) -- see tests/test_oracle_c.py.
 *
 * The arithmetic is written out in exactly the order the reference evaluates it
 * on NumPy (np.convolve accumulates over ascending memory index; python sum()
 * for the smoothness indicators; left-to-right elementwise expressions).  Compile
 * with -ffp-contract=off so that no FMA is formed.  This is also the bit-exact
 * target of the PSK_MATH_STRICT CUDA kernels.
 *
 * It shares the descriptor struct of include/psk.h; all pointers are HOST pointers.
 * Reference citations are relative to /root/reference/src/pyshocks.
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/psk.h"

#define PSO_API __attribute__((visibility("default")))

static inline int nx_of(const psk_desc *d) { return d->n + 2 * d->g; }

/* zero padding of jnp.convolve(..., "same") (convolve.py:113-114) */
static inline double at0(const double *w, int i, int nx) {
  return (i < 0 || i >= nx) ? 0.0 : w[i];
}

/* ------------------------------------------------------------------------- */
/* WENO-JS one-sided values (weno.py:114-157, :166-256; reconstruction.py:351-355) */

/* right-face value of the cell holding c, from (m2, m1, c, p1, p2) = u[i-2..i+2] */
static double weno53_side(double m2, double m1, double c, double p1, double p2, double eps) {
  /* weno.py:218-231, taps listed (i+2, i+1, i, i-1, i-2); np.convolve runs over
     ascending memory index, i.e. from u[i-2] to u[i+2] */
  const double a0 = 13.0 / 12.0, a1 = 1.0 / 4.0;
  double c00 = (m2 * 1.0 + m1 * -2.0) + c * 1.0;
  double c01 = (m2 * 1.0 + m1 * -4.0) + c * 3.0;
  double c10 = (m1 * 1.0 + c * -2.0) + p1 * 1.0;
  double c11 = (m1 * -1.0 + c * 0.0) + p1 * 1.0;
  double c20 = (c * 1.0 + p1 * -2.0) + p2 * 1.0;
  double c21 = (c * 3.0 + p1 * -4.0) + p2 * 1.0;
  /* weno.py:134-140: sum_j a[j] * conv**2 */
  double b0 = a0 * (c00 * c00) + a1 * (c01 * c01);
  double b1 = a0 * (c10 * c10) + a1 * (c11 * c11);
  double b2 = a0 * (c20 * c20) + a1 * (c21 * c21);
  /* weno.py:234-241 interpolation taps */
  double q0 = (m2 * (2.0 / 6.0) + m1 * (-7.0 / 6.0)) + c * (11.0 / 6.0);
  double q1 = (m1 * (-1.0 / 6.0) + c * (5.0 / 6.0)) + p1 * (2.0 / 6.0);
  double q2 = (c * (2.0 / 6.0) + p1 * (5.0 / 6.0)) + p2 * (-1.0 / 6.0);
  /* weno.py:253-256 */
  double e0 = eps + b0, e1 = eps + b1, e2 = eps + b2;
  double al0 = (1.0 / 10.0) / (e0 * e0);
  double al1 = (6.0 / 10.0) / (e1 * e1);
  double al2 = (3.0 / 10.0) / (e2 * e2);
  double tot = (al0 + al1) + al2;
  /* reconstruction.py:355 */
  return ((al0 / tot) * q0 + (al1 / tot) * q1) + (al2 / tot) * q2;
}

/* JS-3: (m1, c, p1) = u[i-1..i+1]; weno.py:166-203 */
static double weno32_side(double m1, double c, double p1, double eps) {
  double c0 = m1 * -1.0 + c * 1.0;
  double c1 = (m1 * 0.0 + c * -1.0) + p1 * 1.0;
  double b0 = 1.0 * (c0 * c0);
  double b1 = 1.0 * (c1 * c1);
  double q0 = m1 * (-1.0 / 2.0) + c * (3.0 / 2.0);
  double q1 = (m1 * 0.0 + c * (1.0 / 2.0)) + p1 * (1.0 / 2.0);
  double e0 = eps + b0, e1 = eps + b1;
  double al0 = (1.0 / 3.0) / (e0 * e0);
  double al1 = (2.0 / 3.0) / (e1 * e1);
  double tot = al0 + al1;
  return (al0 / tot) * q0 + (al1 / tot) * q1;
}

/* ESWENO32 (weno.py:284-296 on the JS-3 stencils): alpha_k = d_k (1 + tau / (eps + beta_k)),
   tau = (u[i+1] - 2 u[i] + u[i-1])^2, zero at the two ends of the array (jnp.pad).
   Returns the right-face value; *om0 (optional) receives omega_0 (weno.py:296, row 0). */
static double esweno32_side(double m1, double c, double p1, double eps, int tau_zero, double *om0) {
  double c0 = m1 * -1.0 + c * 1.0;
  double c1 = (m1 * 0.0 + c * -1.0) + p1 * 1.0;
  double b0 = 1.0 * (c0 * c0);
  double b1 = 1.0 * (c1 * c1);
  double q0 = m1 * (-1.0 / 2.0) + c * (3.0 / 2.0);
  double q1 = (m1 * 0.0 + c * (1.0 / 2.0)) + p1 * (1.0 / 2.0);
  double tt = (p1 - 2.0 * c) + m1; /* u[2:] - 2 * u[1:-1] + u[:-2] */
  double tau = tau_zero ? 0.0 : tt * tt;
  double al0 = (1.0 / 3.0) * (1.0 + tau / (eps + b0));
  double al1 = (2.0 / 3.0) * (1.0 + tau / (eps + b1));
  double tot = al0 + al1;
  if (om0) *om0 = al0 / tot;
  return (al0 / tot) * q0 + (al1 / tot) * q1;
}

/* (fl[i], fr[i]) of reconstruct() at cell i of a zero-padded array
   (reconstruction.py:153-163, :358-377: left value = right value of the reversed array) */
static void reconstruct_cell(int rec, double eps, const double *w, int i, int nx, double *fl,
                             double *fr) {
  if (rec == PSK_REC_CONSTANT) {
    *fl = w[i];
    *fr = w[i];
  } else if (rec == PSK_REC_ESWENO32) {
    /* reconstruction.py:436-437; on the reversed array tau reads (m1 - 2 c) + p1 */
    double m1 = at0(w, i - 1, nx), c = w[i], p1 = at0(w, i + 1, nx);
    int edge = (i == 0 || i == nx - 1);
    *fr = esweno32_side(m1, c, p1, eps, edge, NULL);
    *fl = esweno32_side(p1, c, m1, eps, edge, NULL);
  } else if (rec == PSK_REC_WENOJS32) {
    double m1 = at0(w, i - 1, nx), c = w[i], p1 = at0(w, i + 1, nx);
    *fr = weno32_side(m1, c, p1, eps);
    *fl = weno32_side(p1, c, m1, eps);
  } else {
    double m2 = at0(w, i - 2, nx), m1 = at0(w, i - 1, nx), c = w[i];
    double p1 = at0(w, i + 1, nx), p2 = at0(w, i + 2, nx);
    *fr = weno53_side(m2, m1, c, p1, p2, eps);
    *fl = weno53_side(p2, p1, c, m1, m2, eps);
  }
}

PSO_API int pso_reconstruct(const psk_desc *d, const double *f, double *fl, double *fr) {
  const int nx = nx_of(d);
  for (int r = 0; r < d->batch; ++r) {
    const double *row = f + (size_t)r * d->ld;
    for (int i = 0; i < nx; ++i)
      reconstruct_cell(d->rec, d->eps, row, i, nx, fl + (size_t)r * d->ld + i,
                       fr + (size_t)r * d->ld + i);
  }
  return PSK_OK;
}

/* ------------------------------------------------------------------------- */
/* boundary conditions (scalar.py:418-427, :472-500, :529-540) */

static void apply_boundary_row(const psk_desc *d, int r, const double *u, double *w) {
  const int g = d->g, nx = nx_of(d);
  if (w != u) memcpy(w, u, sizeof(double) * (size_t)nx);
  const double *gh = d->ghost ? d->ghost + (size_t)r * d->ghost_ld : NULL;
  switch (d->bc) {
  case PSK_BC_PERIODIC:
    for (int k = 0; k < g; ++k) w[nx - g + k] = w[g + k];
    for (int k = 0; k < g; ++k) w[k] = w[nx - 2 * g + k];
    break;
  case PSK_BC_DIRICHLET:
    for (int k = 0; k < g; ++k) w[k] = gh[k];
    for (int k = 0; k < g; ++k) w[nx - g + k] = gh[g + k];
    break;
  case PSK_BC_NEUMANN:
    /* ghost k (memory order) mirrors interior cell 2g-1-k on the left and
       2(nx-g)-1-(nx-g+k) on the right; gh holds side*(x[ifrom]-x[ito])*g(t) */
    for (int k = 0; k < g; ++k) w[k] = w[2 * g - 1 - k] + gh[k];
    for (int k = 0; k < g; ++k) w[nx - g + k] = w[nx - g - 1 - k] + gh[g + k];
    break;
  default:
    break;
  }
}

PSO_API int pso_apply_boundary(const psk_desc *d, const double *u, double *w) {
  for (int r = 0; r < d->batch; ++r)
    apply_boundary_row(d, r, u + (size_t)r * d->ld, w + (size_t)r * d->ld);
  return PSK_OK;
}

/* ------------------------------------------------------------------------- */
/* numerical flux at face j+1/2 between cells j and j+1 */

static inline double burgers_flux(double u) { return (u * u) / 2.0; } /* burgers/schemes.py:39 */
/* jnp.maximum / jnp.minimum propagate NaN */
static inline double max_nan(double x, double y) { return (x > y || x != x) ? x : y; }
static inline double min_nan(double x, double y) { return (x < y || x != x) ? x : y; }

typedef struct {
  const double *w;
  const double *ul, *ur; /* reconstruct(w) */
  int nx;
  double lf_speed; /* global max |w| (scalar.py:277) */
} row_ctx;

static double face_flux(const psk_desc *d, const row_ctx *c, int j) {
  const double *w = c->w;
  /* reconstruct() returns whole arrays (reconstruction.py:377); the face j+1/2 uses the
     right value of cell j and the left value of cell j+1 */
  const double urj = c->ur[j], ulp = c->ul[j + 1];
  if (d->equation == PSK_EQ_BURGERS) {
    switch (d->flux) {
    case PSK_FLUX_RUSANOV:
    case PSK_FLUX_LAX_FRIEDRICHS: {
      /* scalar.py:231-249, :277-278 */
      double fr = burgers_flux(urj), fl = burgers_flux(ulp);
      double a = (d->flux == PSK_FLUX_LAX_FRIEDRICHS) ? c->lf_speed
                                                      : max_nan(fabs(w[j + 1]), fabs(w[j]));
      double nu = d->nu ? d->nu[j] : 1.0;
      return 0.5 * (fl + fr) - ((0.5 * a) * nu) * (ulp - urj);
    }
    case PSK_FLUX_UPWIND:
    case PSK_FLUX_ESWENO: {
      /* scalar.py:123-132 with a = u (burgers/schemes.py:89, :255) */
      double aavg = (urj + ulp) / 2.0;
      double fnum = aavg > 0.0 ? burgers_flux(urj) : burgers_flux(ulp);
      if (d->flux == PSK_FLUX_UPWIND) return fnum;
      /* burgers/schemes.py:237-256: omega_0 of cells j, j+1; mu (Equation 37 of Yamaleev2009) */
      double omj, omp;
      esweno32_side(at0(w, j - 1, c->nx), w[j], w[j + 1], d->eps, j == 0, &omj);
      esweno32_side(w[j], w[j + 1], at0(w, j + 2, c->nx), d->eps, j + 1 == c->nx - 1, &omp);
      double dom = omp - omj;
      double mu = sqrt(dom * dom + d->delta * d->delta) / 8.0;
      double gnum = (-(mu + dom / 8.0)) * (w[j + 1] - w[j]);
      return fnum + gnum;
    }
    case PSK_FLUX_ENGQUIST_OSHER: {
      /* scalar.py:311-322 with omega = 0 (burgers/schemes.py:184) */
      double fr = burgers_flux(max_nan(urj, 0.0));
      double fl = burgers_flux(min_nan(ulp, 0.0));
      return (fr + fl) - burgers_flux(0.0);
    }
    }
    return NAN;
  }
  /* advection/schemes.py:100-114, continuity/schemes.py:93-110 */
  double aavg = (d->vel_r[j] + d->vel_l[j + 1]) / 2.0;
  if (d->equation == PSK_EQ_ADVECTION) return aavg > 0.0 ? urj : ulp;
  return aavg > 0.0 ? d->vel_r[j] * urj : d->vel_l[j + 1] * ulp;
}

static double max_abs_range(const double *w, int lo, int hi) {
  double m = 0.0;
  int seen_nan = 0;
  for (int i = lo; i < hi; ++i) {
    double a = fabs(w[i]);
    if (a != a) seen_nan = 1;
    if (a > m) m = a;
  }
  return seen_nan ? NAN : m;
}

/* F[0..nx]: jnp.pad(fnum, 1); lr: scratch of 2 nx doubles for the reconstructed arrays */
static void flux_row(const psk_desc *d, const double *w, double *F, double *lr) {
  const int nx = nx_of(d);
  double *ul = lr, *ur = lr + nx;
  for (int i = 0; i < nx; ++i) reconstruct_cell(d->rec, d->eps, w, i, nx, ul + i, ur + i);
  row_ctx c = {w, ul, ur, nx, 0.0};
  if (d->flux == PSK_FLUX_LAX_FRIEDRICHS && d->equation == PSK_EQ_BURGERS)
    c.lf_speed = max_abs_range(w, 0, nx);
  F[0] = 0.0;
  for (int j = 0; j < nx - 1; ++j) F[j + 1] = face_flux(d, &c, j);
  F[nx] = 0.0;
}

PSO_API int pso_numerical_flux(const psk_desc *d, const double *w, double *F, int64_t ld_f) {
  double *lr = (double *)malloc(sizeof(double) * 2 * (size_t)nx_of(d));
  for (int r = 0; r < d->batch; ++r) flux_row(d, w + (size_t)r * d->ld, F + (size_t)r * ld_f, lr);
  free(lr);
  return PSK_OK;
}

/* schemes.py:339-346, advection/schemes.py:62-73; scratch: w[nx], F[nx+1 .. 3nx+1] */
static void rhs_row(const psk_desc *d, int r, const double *u, double *L, double *w, double *F) {
  const int nx = nx_of(d);
  apply_boundary_row(d, r, u, w);
  flux_row(d, w, F, F + nx + 1);
  if (d->equation == PSK_EQ_ADVECTION) {
    for (int i = 0; i < nx; ++i) L[i] = ((-d->velocity[i]) * (F[i + 1] - F[i])) / d->dx;
  } else {
    for (int i = 0; i < nx; ++i) L[i] = (-(F[i + 1] - F[i])) / d->dx;
  }
}

/* OpenMP thread count of the row loops (bench.py sets it to the affinity core count explicitly:
   torch.distributed.run exports OMP_NUM_THREADS=1 to its workers).  Returns the count in use. */
PSO_API int pso_set_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  return omp_get_max_threads();
#else
  (void)nthreads;
  return 1;
#endif
}

PSO_API int pso_apply_operator(const psk_desc *d, const double *u, double *L) {
  const int nx = nx_of(d);
#pragma omp parallel
  {
    double *w = (double *)malloc(sizeof(double) * (size_t)(4 * nx + 1));
    double *F = w + nx;
#pragma omp for schedule(static)
    for (int r = 0; r < d->batch; ++r)
      rhs_row(d, r, u + (size_t)r * d->ld, L + (size_t)r * d->ld, w, F);
    free(w);
  }
  return PSK_OK;
}

PSO_API int pso_max_abs(const psk_desc *d, const double *u, int interior_only, double *out) {
  const int nx = nx_of(d);
  for (int r = 0; r < d->batch; ++r)
    out[r] = interior_only ? max_abs_range(u + (size_t)r * d->ld, d->g, nx - d->g)
                           : max_abs_range(u + (size_t)r * d->ld, 0, nx);
  return PSK_OK;
}

/* ------------------------------------------------------------------------- */
/* SSPRK33 (timestepping.py:312-320).  ghost3: Dirichlet / Neumann data for the
   three stage times t, t + dt, t + dt/2 as three consecutive blocks of
   (batch or 1) x 2g doubles, or NULL to use d->ghost for all three stages. */

static void step_row(const psk_desc *d0, int r, const double *ghost3, int64_t ghost_block,
                     const double *u, double dt, double *out, double *scratch) {
  const int nx = nx_of(d0);
  double *k1 = scratch, *k2 = scratch + nx, *L = scratch + 2 * nx, *w = scratch + 3 * nx;
  double *F = scratch + 4 * nx;
  psk_desc d = *d0;
  if (ghost3) d.ghost = ghost3;
  rhs_row(&d, r, u, L, w, F);
  for (int i = 0; i < nx; ++i) k1[i] = u[i] + dt * L[i];
  if (ghost3) d.ghost = ghost3 + ghost_block;
  rhs_row(&d, r, k1, L, w, F);
  for (int i = 0; i < nx; ++i) k2[i] = (3.0 / 4.0) * u[i] + (1.0 / 4.0) * (k1[i] + dt * L[i]);
  if (ghost3) d.ghost = ghost3 + 2 * ghost_block;
  rhs_row(&d, r, k2, L, w, F);
  for (int i = 0; i < nx; ++i) out[i] = (1.0 / 3.0) * u[i] + (2.0 / 3.0) * (k2[i] + dt * L[i]);
}

PSO_API int pso_ssprk33_step(const psk_desc *d, const double *u, const double *dt,
                             int64_t dt_stride, const double *ghost3, double *out) {
  const int nx = nx_of(d);
  const int64_t ghost_block = (d->ghost_ld ? (int64_t)d->batch * d->ghost_ld : 2 * d->g);
#pragma omp parallel
  {
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(7 * nx + 1));
#pragma omp for schedule(static)
    for (int r = 0; r < d->batch; ++r)
      step_row(d, r, ghost3, ghost_block, u + (size_t)r * d->ld, dt[(size_t)r * dt_stride],
               out + (size_t)r * d->ld, scratch);
    free(scratch);
  }
  return PSK_OK;
}

/* nsteps fixed-dt steps in place (periodic / ghost data constant in time): the CPU
   baseline workload of bench.py.  Rows are independent, one OpenMP thread each. */
PSO_API int pso_solve_fixed_dt(const psk_desc *d, double *u, double dt, int nsteps) {
  const int nx = nx_of(d);
#pragma omp parallel
  {
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(8 * nx + 1));
    double *tmp = scratch + 7 * nx + 1;
#pragma omp for schedule(static)
    for (int r = 0; r < d->batch; ++r) {
      double *row = u + (size_t)r * d->ld;
      for (int s = 0; s < nsteps; ++s) {
        step_row(d, r, NULL, 0, row, dt, tmp, scratch);
        memcpy(row, tmp, sizeof(double) * (size_t)nx);
      }
    }
    free(scratch);
  }
  return PSK_OK;
}

/* Adaptive time loop of timestepping.step (timestepping.py:128-150) for a Burgers
   scheme on one row with a time-independent boundary: returns the number of steps,
   writes the dt history (up to max_steps) and leaves u(tfinal) in u. */
PSO_API int pso_solve_adaptive(const psk_desc *d, double *u, double theta, double cfl_scale,
                               double tfinal, int max_steps, double *dt_hist) {
  const int nx = nx_of(d);
  double *scratch = (double *)malloc(sizeof(double) * (size_t)(8 * nx + 1));
  double *tmp = scratch + 7 * nx + 1;
  double t = 0.0;
  int m = 0;
  while (!(t >= tfinal) && m < max_steps) {
    double smax = max_abs_range(u, d->g, nx - d->g);
    double dt = theta * (cfl_scale / smax); /* examples/burgers.py:161-162 */
    double dt_min = tfinal - t;
    dt = (dt < dt_min ? dt : dt_min) + 1.0e-15;
    if (!isfinite(dt)) {
      free(scratch);
      return -PSK_E_NONFINITE;
    }
    step_row(d, 0, NULL, 0, u, dt, tmp, scratch);
    memcpy(u, tmp, sizeof(double) * (size_t)nx);
    if (dt_hist) dt_hist[m] = dt;
    m += 1;
    t += dt;
  }
  free(scratch);
  return m;
}
