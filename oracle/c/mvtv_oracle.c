/* CPU ORACLE (test infrastructure, NOT product code) -- see mvtv_oracle.h.
 *
 * Matrix-free C restatement of:
 *   operators : cpp-code/utils.cpp:40-71 (index maps), :73-101 (masks), :103-169 (one-axis forward
 *               difference, +1 at ind / -1 at ind+e), :171-234 (mixed partials incl. the direction-0
 *               quirk at :187), :245-269 (stack order and delta scaling), :311-352 (nearest / O)
 *   solver    : cpp-code/solvers.cpp:15-29,70-88,90-130 (CPP mode),
 *               rcpp-code/MultivarTV/src/solvers.cpp:77-136 (RCPP mode), code/solvers.py:54-76 (PY mode)
 * The x-update (arma::spsolve -> SuperLU in the reference, cpp-code/solvers.cpp:116) is a direct
 * banded Cholesky for small N or a Jacobi-PCG driven to cg_rtol for large N.
 */
#include "mvtv_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------ operator table */
int ora_op_init(ora_op *op, int p, const int64_t *m, const double *deltas, int variant) {
  if (p < 1 || p > ORA_MAXP) return -2;
  memset(op, 0, sizeof(*op));
  op->p = p;
  op->K = (1 << p) - 1;
  int64_t s = 1;
  for (int a = 0; a < p; ++a) {
    op->m[a] = m[a];
    op->stride[a] = s;
    s *= m[a];
  }
  op->N = s;
  int64_t off = 0;
  for (int b = 0; b < op->K; ++b) {
    /* stack order (cpp-code/utils.cpp:258-267): all-ones mask first, then binaries of 1..K-1 */
    int num = (b == 0) ? op->K : b;
    int S = 0; /* bit a set <=> binary[a]==1; binary is MSB-first so axis a <-> bit (p-1-a) of num */
    for (int a = 0; a < p; ++a)
      if ((num >> (p - 1 - a)) & 1) S |= 1 << a;
    double sc = 1.0;
    if (b != 0 && deltas)
      for (int a = 0; a < p; ++a)
        if (!((S >> a) & 1)) sc *= deltas[a];
    int Sp = S;
    if (__builtin_popcount(S) > 1 && variant == ORA_VARIANT_REFERENCE) {
      int lowest = __builtin_ctz(S);
      Sp = (S & ~(1 << lowest)) | 1; /* first factor along axis 0 (cpp-code/utils.cpp:187) */
      if (lowest != 0) {
        /* the sparse product only conforms when m[0]==m[lowest] ... in general when removing one
           from axis 0 vs from axis `lowest` gives the same element count: m[0]==m[lowest] */
        if (m[0] != m[lowest]) return -1;
      }
    }
    op->mask[b] = Sp;
    op->scale[b] = sc;
    int64_t rs = 1;
    for (int a = 0; a < p; ++a) {
      op->rstride[b][a] = rs;
      rs *= m[a] - ((Sp >> a) & 1);
    }
    op->rows[b] = rs;
    op->row_off[b] = off;
    off += rs;
  }
  op->R = off;
  return 0;
}

int64_t ora_rows(int p, const int64_t *m, int variant) {
  ora_op op;
  if (ora_op_init(&op, p, m, NULL, variant) != 0) return -1;
  return op.R;
}

void ora_D_apply(const ora_op *op, const double *theta, double *out) {
  const int p = op->p;
  for (int b = 0; b < op->K; ++b) {
    const int S = op->mask[b];
    const double sc = op->scale[b];
    int64_t rd[ORA_MAXP];
    for (int a = 0; a < p; ++a) rd[a] = op->m[a] - ((S >> a) & 1);
    double *o = out + op->row_off[b];
    const int64_t rows = op->rows[b];
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
      int64_t rem = r, v = 0;
      for (int a = 0; a < p; ++a) {
        int64_t ia = rem % rd[a];
        rem /= rd[a];
        v += ia * op->stride[a];
      }
      double acc = 0.0;
      for (int e = S;; e = (e - 1) & S) { /* all subsets of S, including 0 */
        int64_t vv = v;
        for (int a = 0; a < p; ++a)
          if ((e >> a) & 1) vv += op->stride[a];
        acc += (__builtin_popcount(e) & 1) ? -theta[vv] : theta[vv];
        if (e == 0) break;
      }
      o[r] = sc * acc;
    }
  }
}

void ora_Dt_apply(const ora_op *op, const double *w, double *out) {
  const int p = op->p;
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < op->N; ++v) {
    int64_t idx[ORA_MAXP];
    int64_t rem = v;
    for (int a = 0; a < p; ++a) {
      idx[a] = rem % op->m[a];
      rem /= op->m[a];
    }
    double total = 0.0;
    for (int b = 0; b < op->K; ++b) {
      const int S = op->mask[b];
      const double *wb = w + op->row_off[b];
      double acc = 0.0;
      for (int e = S;; e = (e - 1) & S) {
        int ok = 1;
        int64_t r = 0;
        for (int a = 0; a < p; ++a) {
          int64_t ia = idx[a] - ((e >> a) & 1);
          if ((S >> a) & 1) {
            if (ia < 0 || ia > op->m[a] - 2) { ok = 0; break; }
          }
          r += ia * op->rstride[b][a];
        }
        if (ok) acc += (__builtin_popcount(e) & 1) ? -wb[r] : wb[r];
        if (e == 0) break;
      }
      total += op->scale[b] * acc;
    }
    out[v] = total;
  }
}

/* ------------------------------------------------------------------ nearest / O */
static int64_t nearest_knot(const double *ax, int64_t m, double x) {
  /* argmin_j (x-ax[j])^2, ties -> lower j; ax ascending */
  int64_t lo = 0, hi = m; /* first j with ax[j] >= x */
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (ax[mid] < x) lo = mid + 1; else hi = mid;
  }
  int64_t best = lo < m ? lo : m - 1;
  double bd = (x - ax[best]) * (x - ax[best]);
  for (int64_t j = best - 1; j >= 0 && j >= best - 2; --j) {
    double d = (x - ax[j]) * (x - ax[j]);
    if (d <= bd) { bd = d; best = j; }
  }
  return best;
}

void ora_nearest(int p, const int64_t *m, const double *axes, int64_t n, const double *data,
                 int64_t *idx) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int64_t v = 0, stride = 1;
    const double *ax = axes;
    for (int a = 0; a < p; ++a) {
      v += nearest_knot(ax, m[a], data[i + (int64_t)a * n]) * stride;
      stride *= m[a];
      ax += m[a];
    }
    idx[i] = v;
  }
}

void ora_nearest_brute(int p, const int64_t *m, const double *axes, int64_t n, const double *data,
                       int64_t *idx) {
  int64_t N = 1;
  for (int a = 0; a < p; ++a) N *= m[a];
  for (int64_t i = 0; i < n; ++i) {
    double best = INFINITY;
    int64_t bj = 0;
    for (int64_t j = 0; j < N; ++j) {
      int64_t rem = j;
      const double *ax = axes;
      double d = 0.0;
      for (int a = 0; a < p; ++a) {
        double t = data[i + (int64_t)a * n] - ax[rem % m[a]];
        d += t * t;
        rem /= m[a];
        ax += m[a];
      }
      if (d < best) { best = d; bj = j; } /* strict: first minimum wins (cpp-code/utils.cpp:319-320) */
    }
    idx[i] = bj;
  }
}

void ora_scatter(int64_t n, const int64_t *idx, const double *y, int64_t N, double *Oty, double *counts) {
  for (int64_t v = 0; v < N; ++v) { Oty[v] = 0.0; counts[v] = 0.0; }
  for (int64_t i = 0; i < n; ++i) { /* increasing i: the order arma's Ot*y accumulates in */
    Oty[idx[i]] += y[i];
    counts[idx[i]] += 1.0;
  }
}

/* ------------------------------------------------------------------ small vector helpers */
static double dot(const double *a, const double *b, int64_t n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* ------------------------------------------------------------------ x-update: M = diag(c) + rho*D^T D */
typedef struct {
  const ora_op *op;
  const double *counts;
  /* band Cholesky */
  int64_t bw;
  double *band;
  double band_rho;
  int band_valid;
  /* pcg work */
  double *diagK, *r, *z, *pvec, *q, *tmpR;
  int64_t inner;
} xsolver;

static void build_diagK(const ora_op *op, double *diagK) {
  const int p = op->p;
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < op->N; ++v) {
    int64_t idx[ORA_MAXP], rem = v;
    for (int a = 0; a < p; ++a) { idx[a] = rem % op->m[a]; rem /= op->m[a]; }
    double d = 0.0;
    for (int b = 0; b < op->K; ++b) {
      double t = op->scale[b] * op->scale[b];
      for (int a = 0; a < p; ++a)
        if ((op->mask[b] >> a) & 1) t *= (double)((idx[a] > 0) + (idx[a] < op->m[a] - 1));
      d += t;
    }
    diagK[v] = d;
  }
}

static int band_factor(xsolver *xs, double rho) {
  const ora_op *op = xs->op;
  const int p = op->p;
  const int64_t N = op->N, bw = xs->bw, ld = bw + 1;
  double *A = xs->band;
  memset(A, 0, sizeof(double) * (size_t)(N * ld));
  for (int64_t v = 0; v < N; ++v) A[v * ld] = xs->counts[v];
  for (int b = 0; b < op->K; ++b) {
    const int S = op->mask[b];
    const double w = rho * op->scale[b] * op->scale[b];
    int64_t rd[ORA_MAXP];
    for (int a = 0; a < p; ++a) rd[a] = op->m[a] - ((S >> a) & 1);
    for (int64_t r = 0; r < op->rows[b]; ++r) {
      int64_t rem = r, v = 0;
      for (int a = 0; a < p; ++a) { v += (rem % rd[a]) * op->stride[a]; rem /= rd[a]; }
      for (int e1 = S;; e1 = (e1 - 1) & S) {
        int64_t v1 = v;
        for (int a = 0; a < p; ++a) if ((e1 >> a) & 1) v1 += op->stride[a];
        double s1 = (__builtin_popcount(e1) & 1) ? -1.0 : 1.0;
        for (int e2 = S;; e2 = (e2 - 1) & S) {
          int64_t v2 = v;
          for (int a = 0; a < p; ++a) if ((e2 >> a) & 1) v2 += op->stride[a];
          if (v1 >= v2) {
            double s2 = (__builtin_popcount(e2) & 1) ? -1.0 : 1.0;
            A[v2 * ld + (v1 - v2)] += w * s1 * s2;
          }
          if (e2 == 0) break;
        }
        if (e1 == 0) break;
      }
    }
  }
  for (int64_t j = 0; j < N; ++j) {
    double d = A[j * ld];
    if (!(d > 0.0)) return -1;
    d = sqrt(d);
    A[j * ld] = d;
    int64_t kmax = (N - 1 - j < bw) ? N - 1 - j : bw;
    for (int64_t k = 1; k <= kmax; ++k) A[j * ld + k] /= d;
    for (int64_t k = 1; k <= kmax; ++k) {
      const double ljk = A[j * ld + k];
      if (ljk == 0.0) continue;
      double *col = A + (j + k) * ld;
      for (int64_t l = k; l <= kmax; ++l) col[l - k] -= ljk * A[j * ld + l];
    }
  }
  xs->band_rho = rho;
  xs->band_valid = 1;
  return 0;
}

static void band_solve(const xsolver *xs, const double *b, double *x) {
  const int64_t N = xs->op->N, bw = xs->bw, ld = bw + 1;
  const double *A = xs->band;
  for (int64_t j = 0; j < N; ++j) x[j] = b[j];
  for (int64_t j = 0; j < N; ++j) {
    x[j] /= A[j * ld];
    int64_t kmax = (N - 1 - j < bw) ? N - 1 - j : bw;
    const double xj = x[j];
    for (int64_t k = 1; k <= kmax; ++k) x[j + k] -= A[j * ld + k] * xj;
  }
  for (int64_t j = N - 1; j >= 0; --j) {
    int64_t kmax = (N - 1 - j < bw) ? N - 1 - j : bw;
    double s = x[j];
    for (int64_t k = 1; k <= kmax; ++k) s -= A[j * ld + k] * x[j + k];
    x[j] = s / A[j * ld];
  }
}

static void apply_M(xsolver *xs, double rho, const double *x, double *out) {
  const ora_op *op = xs->op;
  ora_D_apply(op, x, xs->tmpR);
  ora_Dt_apply(op, xs->tmpR, out);
  const double *c = xs->counts;
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < op->N; ++v) out[v] = c[v] * x[v] + rho * out[v];
}

static int pcg_solve(xsolver *xs, double rho, const double *b, double *x, double rtol, int maxit) {
  const int64_t N = xs->op->N;
  double *r = xs->r, *z = xs->z, *pv = xs->pvec, *q = xs->q;
  const double *c = xs->counts, *dK = xs->diagK;
  apply_M(xs, rho, x, q);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) r[i] = b[i] - q[i];
  const double bnorm = sqrt(dot(b, b, N));
  const double thresh = rtol * bnorm;
  double rr = dot(r, r, N);
  if (sqrt(rr) <= thresh) return 0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) { z[i] = r[i] / (c[i] + rho * dK[i]); pv[i] = z[i]; }
  double rz = dot(r, z, N);
  int it = 0;
  while (it < maxit) {
    apply_M(xs, rho, pv, q);
    const double alpha = rz / dot(pv, q, N);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) { x[i] += alpha * pv[i]; r[i] -= alpha * q[i]; }
    ++it;
    rr = dot(r, r, N);
    if (sqrt(rr) <= thresh) break;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) z[i] = r[i] / (c[i] + rho * dK[i]);
    const double rz_new = dot(r, z, N);
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) pv[i] = z[i] + beta * pv[i];
  }
  return it;
}

static int x_solve(xsolver *xs, const ora_params *prm, double rho_m, const double *b, double *theta) {
  if (prm->solver == ORA_SOLVER_BANDCHOL) {
    if (!xs->band_valid || xs->band_rho != rho_m)
      if (band_factor(xs, rho_m) != 0) return -1;
    band_solve(xs, b, theta);
    return 0;
  }
  double rtol = prm->cg_rtol > 0 ? prm->cg_rtol : 1e-13;
  int maxit = prm->cg_maxit > 0 ? prm->cg_maxit : 100000;
  xs->inner += pcg_solve(xs, rho_m, b, theta, rtol, maxit);
  return 0;
}

/* ------------------------------------------------------------------ the ADMM loops */
static inline double soft(double z, double kappa) {
  /* cpp-code/solvers.cpp:24-29: sign(z) * max(|z| - kappa, 0); kappa may be +inf */
  double mag = fabs(z) - kappa;
  if (!(mag > 0.0)) mag = 0.0;
  double sg = (z > 0.0) - (z < 0.0);
  return sg * mag;
}

int ora_admm(const ora_op *op, const double *Oty, const double *counts, double mean_y,
             const ora_params *prm, const double *theta_init, const double *u_init,
             double *theta, double *u_out, ora_result *res) {
#ifdef _OPENMP
  if (prm->nthreads > 0) omp_set_num_threads(prm->nthreads);
#endif
  const int64_t N = op->N, R = op->R;
  const int mode = prm->mode;
  const double lambda = prm->lambda;
  double tol = prm->tol > 0 ? prm->tol : (mode == ORA_MODE_RCPP ? 1e-4 : 1e-3);
  int max_counter = prm->max_counter > 0 ? prm->max_counter
                    : (mode == ORA_MODE_CPP ? 2000 : mode == ORA_MODE_RCPP ? 3000 : 5000);
  memset(res, 0, sizeof(*res));

  xsolver xs;
  memset(&xs, 0, sizeof(xs));
  xs.op = op;
  xs.counts = counts;
  if (prm->solver == ORA_SOLVER_BANDCHOL) {
    int64_t bw = 0;
    for (int b = 0; b < op->K; ++b) {
      int64_t t = 0;
      for (int a = 0; a < op->p; ++a) if ((op->mask[b] >> a) & 1) t += op->stride[a];
      if (t > bw) bw = t;
    }
    xs.bw = bw;
    xs.band = (double *)malloc(sizeof(double) * (size_t)(N * (bw + 1)));
    if (!xs.band) return -3;
  } else {
    xs.diagK = (double *)malloc(sizeof(double) * (size_t)N);
    xs.r = (double *)malloc(sizeof(double) * (size_t)N);
    xs.z = (double *)malloc(sizeof(double) * (size_t)N);
    xs.pvec = (double *)malloc(sizeof(double) * (size_t)N);
    xs.q = (double *)malloc(sizeof(double) * (size_t)N);
    xs.tmpR = (double *)malloc(sizeof(double) * (size_t)R);
    build_diagK(op, xs.diagK);
  }
  double *alpha = (double *)malloc(sizeof(double) * (size_t)R);
  double *u = (double *)malloc(sizeof(double) * (size_t)R);
  double *w = (double *)malloc(sizeof(double) * (size_t)R);      /* scratch rows */
  double *Dth = (double *)malloc(sizeof(double) * (size_t)R);
  double *prim = (double *)malloc(sizeof(double) * (size_t)R);
  double *b = (double *)malloc(sizeof(double) * (size_t)N);
  double *tN = (double *)malloc(sizeof(double) * (size_t)N);
  double *thetaold = (double *)malloc(sizeof(double) * (size_t)N);

  /* initial state: cpp-code/solvers.cpp:92-108 ; rcpp solvers.cpp:98-109 ; code/solvers.py:54-65 */
  for (int64_t i = 0; i < N; ++i) theta[i] = theta_init ? theta_init[i] : mean_y;
  ora_D_apply(op, theta, alpha);
  double rho;   /* CPP keeps an int in a double: truncation is applied explicitly below */
  double rho_m = prm->rho_matrix0;
  if (mode == ORA_MODE_CPP) {
    for (int64_t i = 0; i < R; ++i) u[i] = 1.0 / lambda;
    for (int64_t i = 0; i < N; ++i) thetaold[i] = mean_y - 0.1;
    rho = (double)(int)lambda; /* int rho = lambda;  cpp-code/solvers.cpp:108 */
  } else if (mode == ORA_MODE_PY) {
    for (int64_t i = 0; i < R; ++i) u[i] = 1.0 / lambda;
    for (int64_t i = 0; i < N; ++i) thetaold[i] = mean_y - 1.0;
    rho = lambda;
  } else {
    for (int64_t i = 0; i < R; ++i) u[i] = u_init ? u_init[i] : 0.0;
    rho = prm->rho_init;
  }
  int counter = 1, passes = 0, status = 0;
  double dual_norm = 1.0, primal_norm = 1.0, eps_dual = tol, eps_primal = tol;
  double r_norm = NAN, s_norm = NAN;
  const double t0 = now_s();
  for (;;) {
    /* loop test */
    if (mode == ORA_MODE_RCPP) {
      if (!(dual_norm > eps_dual || primal_norm > eps_primal)) break;       /* rcpp :110 */
    } else {
      int any = 0;
#pragma omp parallel for reduction(| : any) schedule(static)
      for (int64_t i = 0; i < N; ++i) any |= (fabs(theta[i] - thetaold[i]) > tol); /* cpp :113 */
      if (!any) break;
    }
    if (prm->max_passes > 0 && passes >= prm->max_passes) break;
    memcpy(thetaold, theta, sizeof(double) * (size_t)N);
    /* b = Oty + rho * Dt*(alpha+u)   cpp :115 / rcpp :112 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < R; ++i) w[i] = alpha[i] + u[i];
    ora_Dt_apply(op, w, tN);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) b[i] = Oty[i] + rho * tN[i];
    /* theta = spsolve(sp_crosses, b)  cpp :116 / rcpp :113 */
    if (x_solve(&xs, prm, rho_m, b, theta) != 0) { status = -1; break; }
    /* alpha = softthresh(D*theta - u, lambda/rho)  cpp :117 / rcpp :114 */
    const double kappa = (rho != 0.0) ? lambda / rho : INFINITY;
    ora_D_apply(op, theta, Dth);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < R; ++i) alpha[i] = soft(Dth[i] - u[i], kappa);
    if (mode == ORA_MODE_CPP) {
      /* dual_residual = rho*Dt*(alpha+u) with u BEFORE its update   cpp :118 */
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < R; ++i) w[i] = alpha[i] + u[i];
      ora_Dt_apply(op, w, tN);
      s_norm = fabs(rho) * sqrt(dot(tN, tN, N));
    }
    /* primal_residual = alpha - D*theta ; u += primal_residual   cpp :119-120 / rcpp :115-116 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < R; ++i) {
      prim[i] = alpha[i] - Dth[i];
      w[i] = prim[i]; /* u_new - u_old, for the RCPP dual residual */
      u[i] += prim[i];
    }
    r_norm = sqrt(dot(prim, prim, R));
    ++passes;
    if (mode == ORA_MODE_PY) continue; /* no residuals, no adaptation, counter never moves */
    if (mode == ORA_MODE_CPP) {
      counter += 1;
      if (counter > max_counter) { status = 1; break; } /* throw  cpp :122-124 */
      double rho_next = rho;
      double uscale = 1.0;
      if (r_norm > 20 * s_norm) { rho_next = 20 * rho; uscale = 0.05; }       /* cpp :75-79 */
      else if (s_norm > 20 * r_norm) { rho_next = 0.1 * rho; uscale = 10; }   /* cpp :80-83 */
      if (uscale != 1.0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < R; ++i) u[i] = uscale * u[i];
      }
      rho = (double)(int)rho_next; /* rho = stepobject.rho_next into an int  cpp :126 */
    } else {
      /* dual_residual = rho*Dt*(u - uold)  rcpp :117 */
      ora_Dt_apply(op, w, tN);
      s_norm = fabs(rho) * sqrt(dot(tN, tN, N));
      dual_norm = s_norm;
      primal_norm = r_norm;
      ora_Dt_apply(op, u, tN);
      eps_dual = tol * (sqrt((double)N) + sqrt(dot(tN, tN, N)));                 /* rcpp :121 */
      double nD = sqrt(dot(Dth, Dth, R)), nA = sqrt(dot(alpha, alpha, R));
      eps_primal = tol * (sqrt((double)R) + (nD > nA ? nD : nA));                /* rcpp :122 */
      double uscale = 1.0;
      if (r_norm > 10 * s_norm) { rho = 2.0 * rho; uscale = 1.0 / 2.0; }         /* rcpp :82-85 */
      else if (s_norm > 10 * r_norm) { rho = 1.0 / 2.0 * rho; uscale = 2.0; }    /* rcpp :86-89 */
      if (uscale != 1.0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < R; ++i) u[i] = uscale * u[i];
      }
      rho_m = rho;                                                               /* rcpp :126 */
      counter += 1;
      if (counter > max_counter) { status = 1; break; }                          /* rcpp :129-132 */
    }
  }
  res->seconds = now_s() - t0;
  res->counter = counter;
  res->passes = passes;
  res->status = status;
  res->rho = rho;
  res->r_norm = r_norm;
  res->s_norm = s_norm;
  res->inner_iters = xs.inner;
  if (u_out) memcpy(u_out, u, sizeof(double) * (size_t)R);
  free(alpha); free(u); free(w); free(Dth); free(prim); free(b); free(tN); free(thetaold);
  free(xs.band); free(xs.diagK); free(xs.r); free(xs.z); free(xs.pvec); free(xs.q); free(xs.tmpR);
  return status < 0 ? status : 0;
}
