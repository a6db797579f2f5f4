/* CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the MultivarTV mesh-based ADMM hot path, independent of
 * oracle/py_oracle.py (which materialises D and uses SuperLU).  This one is matrix-free and is the
 * implementation that bench.py times as the CPU baseline ("port").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Pinning status: see the header of oracle/py_oracle.py and DESIGN.md.
 */
#ifndef MVTV_ORACLE_H
#define MVTV_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAXP 6
#define ORA_MAXK 63

#define ORA_MODE_CPP 0  /* cpp-code/solvers.cpp:90-130 */
#define ORA_MODE_RCPP 1 /* rcpp-code/MultivarTV/src/solvers.cpp:96-136 */
#define ORA_MODE_PY 2   /* code/solvers.py:54-76 */

#define ORA_VARIANT_REFERENCE 0 /* mixedpartial first factor along axis 0 (cpp-code/utils.cpp:187) */
#define ORA_VARIANT_INTENDED 1  /* the commented-out intent (cpp-code/utils.cpp:186) */

#define ORA_SOLVER_BANDCHOL 0 /* direct banded Cholesky (small N) */
#define ORA_SOLVER_PCG 1      /* matrix-free Jacobi-PCG */

typedef struct {
  int p, K;
  int64_t m[ORA_MAXP], stride[ORA_MAXP], N, R;
  int mask[ORA_MAXK];        /* effective axis set S' of block b, bit a = difference along axis a */
  double scale[ORA_MAXK];    /* c_S = prod_{k not in S} delta_k (original mask S), 1 for all-ones */
  int64_t rows[ORA_MAXK], row_off[ORA_MAXK];
  int64_t rstride[ORA_MAXK][ORA_MAXP]; /* strides of the reduced (m_a - [a in S']) tensor */
} ora_op;

/* returns 0, or -1 when the reference would fail with a size mismatch (non-cubic p>=3 quirk) */
int ora_op_init(ora_op *op, int p, const int64_t *m, const double *deltas /*nullable*/, int variant);
int64_t ora_rows(int p, const int64_t *m, int variant);
void ora_D_apply(const ora_op *op, const double *theta, double *out /*R*/);
void ora_Dt_apply(const ora_op *op, const double *w /*R*/, double *out /*N*/);

/* nearest vertex, separable form of cpp-code/utils.cpp:311-330 (ties -> lower index) */
void ora_nearest(int p, const int64_t *m, const double *axes_concat, int64_t n,
                 const double *data_colmajor, int64_t *idx);
/* literal O(n*N) brute force of cpp-code/utils.cpp:311-330 for cross-checking */
void ora_nearest_brute(int p, const int64_t *m, const double *axes_concat, int64_t n,
                       const double *data_colmajor, int64_t *idx);
/* Oty = O^T y, counts = diag(O^T O)  (cpp-code/solvers.cpp:37-40) */
void ora_scatter(int64_t n, const int64_t *idx, const double *y, int64_t N, double *Oty, double *counts);

typedef struct {
  int mode, variant, solver;
  double lambda;
  double rho_init;     /* RCPP: caller's rho; ignored by CPP ((int)lambda) and PY (lambda) */
  double rho_matrix0;  /* scalar of the cached system matrix used until the loop rebuilds it */
  double tol;          /* <=0: reference default (CPP/PY 1e-3, RCPP 1e-4) */
  int max_counter;     /* <=0: reference default (2000 / 3000 / 5000) */
  int max_passes;      /* >0: stop after this many passes regardless (bounded benchmark sample) */
  double cg_rtol;      /* PCG: stop at ||b - M x|| <= cg_rtol * ||b|| */
  int cg_maxit;
  int nthreads;        /* <=0: OpenMP default */
} ora_params;

typedef struct {
  int counter;     /* the value the reference prints ("Counter") */
  int passes;      /* loop passes actually executed */
  int status;      /* 0 ok, 1 = counter exceeded max_counter (CPP: the reference throws) */
  double rho, r_norm, s_norm;
  int64_t inner_iters; /* PCG iterations summed over passes */
  double seconds;      /* wall time of the loop */
} ora_result;

int ora_admm(const ora_op *op, const double *Oty, const double *counts, double mean_y,
             const ora_params *prm, const double *theta_init /*nullable*/,
             const double *u_init /*nullable*/, double *theta_out /*N*/, double *u_out /*R, nullable*/,
             ora_result *res);

#ifdef __cplusplus
}
#endif
#endif
