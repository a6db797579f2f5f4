/* mvtv.h -- C ABI of the B200-native MultivarTV hot path (libmvtv_b200.so).
 *
 * Drop-in boundary for the mesh-based ADMM solver of brayano/MultivarTV.  Every entry point cites the
 * reference interface it replaces (paths relative to the upstream repository root).  Plain pointers
 * and sizes only; all pointers are HOST pointers unless the name ends in _dev.  The caller owns every
 * host buffer; a plan owns its device memory, stream and (optional) NCCL communicator.  One plan per
 * host thread.  There is NO CPU fallback: every function returns MVTV_ERR_CUDA if no sm_100 device /
 * driver is usable.
 *
 * Layout conventions (identical to the reference):
 *   data   n x p doubles, COLUMN-major (arma::mat, cpp-code/solvers.hpp:89)
 *   theta  N = prod(m) doubles, mesh flattened column-major, axis 0 fastest (cpp-code/utils.cpp:40-52)
 *   u      R doubles, rows of D in create_D order (cpp-code/utils.cpp:245-269): all-ones mask block first,
 *          then masks 1..K-1 of fd_binaries; inside a block rows follow the column-major enumeration of
 *          the vertices that own a forward difference (cpp-code/utils.cpp:103-127)
 *   axes   concatenated per-axis knot vectors (m[0] + ... + m[p-1] doubles, each ascending); the
 *          tensor-product mesh of create_mesh (cpp-code/utils.cpp:271-298) is axes[k][i_k]
 */
#ifndef MVTV_H
#define MVTV_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVTV_ABI_VERSION 2
#define MVTV_MAXP 4

/* status codes */
#define MVTV_OK 0
#define MVTV_ERR_INVALID 1        /* bad argument */
#define MVTV_ERR_CUDA 2           /* CUDA / NCCL failure, or no usable device (no CPU fallback) */
#define MVTV_ERR_NOT_CONVERGED 3  /* counter exceeded max_counter: CPP mode = the reference's
                                     throw std::invalid_argument("Failed to converge!")
                                     (cpp-code/solvers.cpp:122-124); RCPP mode = message + break
                                     (rcpp solvers.cpp:129-132; outputs are still filled) */
#define MVTV_ERR_DIM_MISMATCH 4   /* reference operator undefined: non-conforming sparse product for a
                                     non-cubic mesh with p>=3 (cpp-code/utils.cpp:187,216) */
#define MVTV_ERR_UNSUPPORTED 5
#define MVTV_ERR_INNER_SOLVE 6    /* x-update did not reach cg_rtol within cg_maxit */

/* solver modes = the three sibling implementations of the same loop */
#define MVTV_MODE_CPP 0   /* cpp-code/solvers.cpp:90-130  (int rho, fixed matrix, |dtheta| stop) */
#define MVTV_MODE_RCPP 1  /* rcpp-code/MultivarTV/src/solvers.cpp:96-136 (double rho, residual stop) */
#define MVTV_MODE_PY 2    /* code/solvers.py:54-76 (fixed rho = lambda, |dtheta| stop) */

/* operator variants */
#define MVTV_VARIANT_REFERENCE 0 /* mixedpartial first factor along axis 0 (cpp-code/utils.cpp:187) */
#define MVTV_VARIANT_INTENDED 1  /* the commented-out intent (cpp-code/utils.cpp:186) */

#define MVTV_F64 64
#define MVTV_F32 32

/* x-update preconditioner for the matrix-free CG that replaces arma::spsolve (cpp-code/solvers.cpp:116) */
#define MVTV_PRECOND_JACOBI 0 /* z = D^-1 r */
#define MVTV_PRECOND_CHEB1 1  /* z = P(D^-1 M) D^-1 r, P = the degree-1 polynomial whose residual is the Chebyshev T2 on
                                 [b/30, b], b = a Gershgorin bound of spec(D^-1 M): one more stencil per CG iteration,
                                 ~1.7x fewer iterations; same solution to cg_rtol */
#define MVTV_PRECOND_AUTO 2   /* Jacobi while the previous x-update on this plan needed <= 24 Jacobi-equivalent iterations,
                                 else the degree measured fastest for the plan's kernels (3 on one GPU with 2-D / 3-D
                                 meshes of even m[0], 1 otherwise) */
#define MVTV_PRECOND_CHEB2 3  /* degree 2..4 of the same Chebyshev family, evaluated in Horner form (one stencil pass per  */
#define MVTV_PRECOND_CHEB3 4  /* degree): ~2.5x / 3.2x / 3.9x fewer iterations than Jacobi on 4096^2.  Plans whose kernels do */
#define MVTV_PRECOND_CHEB4 5  /* not implement the degree (several GPUs, 4-D, odd m[0]) use the highest one they have.      */

/* flags for mvtv_solve */
#define MVTV_WARM_THETA_FROM_PLAN 1u /* theta_init ignored: continue from the theta left on the device */
#define MVTV_WARM_U_FROM_PLAN 2u     /* RCPP: u (and its lazy scale) continue from the device state */

typedef struct mvtv_plan mvtv_plan;

typedef struct {
  int32_t struct_size;   /* sizeof(mvtv_plan_desc) */
  int32_t p;             /* number of covariates / mesh axes, 1..MVTV_MAXP */
  int64_t m[MVTV_MAXP];  /* mesh dims ("vec m", cpp-code/solvers.hpp:89) */
  int32_t dtype;         /* MVTV_F64 | MVTV_F32 (storage and arithmetic of the iteration) */
  int32_t variant;       /* MVTV_VARIANT_* */
  int32_t device;        /* CUDA ordinal, -1 = current device */
  int32_t rank, world;   /* slab partition of the LAST mesh axis; world==1: single GPU */
  int32_t reserved;
  const double *deltas;  /* p mesh widths (create_deltas, cpp-code/utils.cpp:300-307), or NULL:
                            the stand-alone mbs_one path never sets them -> all block scales 1
                            (cpp-code/solvers.cpp:141-145) */
  const void *nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks; NULL when world==1 */
} mvtv_plan_desc;

typedef struct {
  int32_t struct_size;  /* sizeof(mvtv_solve_params) */
  int32_t mode;         /* MVTV_MODE_* */
  double lambda;        /* tuning parameter (cpp-code/solvers.hpp:85,89) */
  double rho_init;      /* RCPP: caller's rho (rcpp solvers.hpp:100); NaN -> lambda/5 (rcpp solvers.cpp:209).
                           Ignored in CPP ((int)lambda, cpp-code/solvers.cpp:108) and PY (lambda). */
  double rho_matrix0;   /* scalar s of the cached system matrix crossO + s*crossD used until the loop
                           rebuilds it; NaN -> lambda (cpp-code/solvers.cpp:144, rcpp solvers.cpp:150) */
  double tol;           /* NaN or <=0 -> the reference's TOL macro (1e-3 cpp/py, 1e-4 rcpp) */
  int32_t max_counter;  /* 0 -> reference default (2000 / 3000 / 5000) */
  int32_t max_passes;   /* >0: stop after this many ADMM passes (fixed-budget benchmarking) */
  double cg_rtol;       /* x-update stops at ||b - M theta|| <= cg_rtol*||b||; <=0 -> 1e-13 (f64) / 1e-5 (f32) */
  int32_t cg_maxit;     /* 0 -> 20000 (f64) / 1000 (f32; reaching it is not an error in f32) */
  int32_t precond;      /* MVTV_PRECOND_* */
  uint32_t flags;       /* MVTV_WARM_* */
  int32_t timing_skip_passes; /* benchmarking: device_seconds and the timed_* results start after this many ADMM passes (the
                           warm-up passes run inside the same call, so no pass is special); 0 = time the whole loop */
} mvtv_solve_params;

typedef struct {
  int32_t counter;      /* the value the reference prints as "Counter" (cpp-code/solvers.cpp:128) */
  int32_t passes;       /* ADMM passes executed (= counter-1 in CPP/RCPP mode) */
  int32_t status;       /* MVTV_OK | MVTV_ERR_NOT_CONVERGED | MVTV_ERR_INNER_SOLVE */
  int32_t reserved;
  double rho;           /* final rho (rcpp admm_out.rho, rcpp solvers.hpp:91-95) */
  double r_norm;        /* last ||primal_residual||_2 */
  double s_norm;        /* last ||dual_residual||_2 */
  double max_dtheta;    /* last max|theta - thetaold| (the CPP/PY loop test, cpp-code/solvers.cpp:113) */
  int64_t inner_iters;  /* CG iterations summed over all passes */
  double device_seconds;/* CUDA-event time of the ADMM loop on the plan's stream (after timing_skip_passes passes) */
  int64_t kernel_launches; /* kernels of this library launched by this call */
  int64_t timed_inner_iters;     /* CG iterations, ADMM passes and kernel launches inside the device_seconds interval */
  int64_t timed_kernel_launches;
  int32_t timed_passes;
  int32_t reserved2;
} mvtv_solve_result;

/* -- library ------------------------------------------------------------------------------------ */
int mvtv_abi_version(void);
const char *mvtv_last_error(void);            /* thread-local message of the last failing call */
int mvtv_device_count(int *count);
/* 128-byte ncclUniqueId for mvtv_plan_desc.nccl_unique_id: call on one rank, broadcast to the others
 * (e.g. with torch.distributed), then every rank creates its plan. */
int mvtv_nccl_unique_id(void *out128);
/* Page-locked (pinned) host buffers.  Not a reference function: the reference never leaves host memory.  Every
 * host pointer of this ABI may be pageable; when it was obtained here, the cudaMemcpyAsync behind the call is a
 * direct DMA transfer at PCIe speed instead of a staged copy (points in, theta / fitted / u out). */
int mvtv_host_alloc(void **out, uint64_t bytes);
int mvtv_host_free(void *ptr);

/* -- plan = operators of create_cache_objects (cpp-code/solvers.cpp:31-41) kept on the device ----- */
/* Replaces: create_D (cpp-code/utils.cpp:245-269) -- D is never materialised, only its block table. */
int mvtv_plan_create(mvtv_plan **plan, const mvtv_plan_desc *desc);
int mvtv_plan_destroy(mvtv_plan *plan);
/* rows of D (inits.rowsD, cpp-code/solvers.cpp:38), vertices N, local slab [z0, z0+nz) of the last axis */
int mvtv_plan_info(const mvtv_plan *plan, int64_t *N, int64_t *R, int64_t *z0, int64_t *nz);

/* Replaces: nearest_interp_matrix + Ot*y + Ot*O (cpp-code/utils.cpp:311-352, solvers.cpp:32,37,40).
 * Bins every point to its nearest vertex (ties -> lowest index), sorts by vertex and reduces segments
 * (deterministic, increasing point index = arma's accumulation order).  With world>1 each rank passes
 * exactly the points whose nearest vertex lies in its slab (multivartv_b200/partition.py buckets and
 * exchanges them); a foreign point is MVTV_ERR_INVALID; mean(y) is all-reduced over the ranks. */
int mvtv_plan_set_points(mvtv_plan *plan, int64_t n, const double *data_colmajor, const double *y,
                         const double *axes);
/* Same, for a dense data matrix in either layout: element (point i, axis a) = data[i*ld_point + a*ld_axis];
 * (1, n) is arma::mat's column-major, (p, 1) is a C / numpy row-major n x p array (no host transpose needed). */
int mvtv_plan_set_points_strided(mvtv_plan *plan, int64_t n, const double *data, int64_t ld_point, int64_t ld_axis,
                                 const double *y, const double *axes);
/* Same with inputs already resident in HBM (device pointers on the plan's device). */
int mvtv_plan_set_points_dev(mvtv_plan *plan, int64_t n, const double *data_colmajor_dev,
                             const double *y_dev, const double *axes_dev);
/* Which kernels this plan runs, as a small JSON object (zu: k_zu_march | k_zu; cg_step: k_cg_step | k_cg_step2d |
 * k_cg_step3d; cg_prec_words: words of HBM traffic per vertex of the first preconditioner pass; fused_update: 1 when
 * k_cg_update and the first preconditioner pass are one kernel; max_degree / auto_degree: polynomial degrees the plan
 * implements / MVTV_PRECOND_AUTO picks; last_degree: degree of the latest x-update; collectives: none | peer | nccl).
 * bench.py uses it for the algorithmic bytes of each kernel class. */
int mvtv_plan_describe(const mvtv_plan *plan, char *buf, int64_t cap);
/* Per-kernel-class CUDA-event timing on the plan's stream (used by bench.py for the roofline).
 * ms[k], count[k], k = MVTV_KC_*: accumulated milliseconds and number of launches since enable. */
#define MVTV_KC_ZU 0        /* fused z/u update + D^T products + norms */
#define MVTV_KC_ZU_INIT 1   /* same kernel, initial D^T D theta / D^T u pass */
#define MVTV_KC_CG_INIT 2   /* b, r = b - M theta, p */
#define MVTV_KC_CG_STEP 3   /* p = z + beta p fused with q = M p, p.q */
#define MVTV_KC_CG_UPDATE 4 /* theta, r update, r.z, r.r (fused_update: + the first preconditioner pass) */
#define MVTV_KC_CG_PREC 5   /* polynomial preconditioner passes not fused into the update: z = P(D^-1 M) D^-1 r, r.z */
#define MVTV_KC_N 8
int mvtv_plan_profile(mvtv_plan *plan, int enable);
int mvtv_plan_get_profile(mvtv_plan *plan, double *ms, int64_t *count);

/* Debug / parity access to the cached operators: Oty (N), diag(crossO) (N), nearest vertex of each point (n) */
int mvtv_plan_get_cache(mvtv_plan *plan, double *Oty, double *counts, int64_t *vertex_of_point);

/* -- the hot path ---------------------------------------------------------------------------------
 * Replaces: admm_update (cpp-code/solvers.hpp:85; rcpp solvers.hpp:100) and the solve+fitted part of
 * mbs_one (cpp-code/solvers.hpp:89).  theta_init: N doubles or NULL (-> mean(y), cpp-code/solvers.cpp:94-99).
 * u_inout: RCPP warm start in / final u out (R doubles, reference row order) or NULL (-> zeros in, nothing out).
 * theta_out: N doubles.  fitted_out: n doubles (O*theta, cpp-code/solvers.cpp:66) or NULL.
 * With world>1, theta_init/theta_out are the rank's LOCAL slab (nz*prod(m[0..p-2]) doubles) and u_inout must be NULL. */
int mvtv_solve(mvtv_plan *plan, const mvtv_solve_params *prm, const double *theta_init, double *u_inout,
               double *theta_out, double *fitted_out, mvtv_solve_result *res);

/* Replaces: mbs_path (cpp-code/solvers.cpp:196-217 ; rcpp solvers.cpp:204-222): the lambdas are solved strictly in
 * order, each warm-started from the previous one WITHOUT leaving the device (CPP/PY mode: theta carried, system
 * matrix crossO + lambda_i*crossD; RCPP mode: theta, u and rho carried, first-pass matrix crossO + rho*crossD).
 * prm->lambda is ignored.  ftrue: n doubles (pass y for the reference default, gen_ftrue solvers.cpp:235-244).
 * mses_out[i] = mse(fitted_i, ftrue) (n_lambda doubles); counters_out, rhos_out (final rho of each solve) and
 * thetas_out (n_lambda x N) are nullable; theta_best_out / fitted_best_out / best_index_out: the first lambda attaining the lowest MSE
 * (fill_output_mbs, solvers.cpp:170-177).  total: passes, inner iterations, device seconds summed over the path.
 * A CPP-mode non-convergence stops the path with MVTV_ERR_NOT_CONVERGED, like the uncaught throw upstream. */
int mvtv_solve_path(mvtv_plan *plan, const mvtv_solve_params *prm, int32_t n_lambda, const double *lambdas,
                    const double *ftrue, double *mses_out, int32_t *counters_out, double *rhos_out,
                    double *thetas_out, double *theta_best_out, double *fitted_best_out, int32_t *best_index_out,
                    mvtv_solve_result *total);

/* Replaces: lam_max_pinv / mypinv / cg (cpp-code/utils.cpp:354-404 for MVTV_MODE_CPP and _PY: CG on D^T D from
 * x0 = mean(Oty), stop at ||r|| < 0.01 or 100 (500 if N < 400) iterations, lambda_max = max|D x|;
 * rcpp-code/MultivarTV/src/utils.cpp:306-355 for MVTV_MODE_RCPP: CGNR, relative 1e-4, min(N,2000) iterations,
 * 5*||D x||_inf).  The CG is truncated, not converged, so the value is only reproducible to ~1e-6 relative across
 * arithmetic orders.  The lambda grid itself (create_lambdas, cpp solvers.cpp:179-192) is host arithmetic. */
int mvtv_lambda_max(mvtv_plan *plan, int mode, double *lambda_max, int32_t *cg_iters);

/* Replaces: mbs_predict (cpp-code/solvers.hpp:93): nearest vertex of each new point, gather theta. */
int mvtv_predict(mvtv_plan *plan, int64_t n_new, const double *data_colmajor, const double *axes,
                 const double *theta, double *fits_out);

/* -- operator-level entry points (parity tests, and drop-in for the reference's free functions) ---- */
/* Replaces: D*theta and Dt*w (cpp-code/solvers.cpp:100,115,117-119) in the reference row order. */
int mvtv_apply_D(mvtv_plan *plan, const double *theta, double *out_rows);
int mvtv_apply_Dt(mvtv_plan *plan, const double *rows, double *out_vertices);
/* Replaces: (crossO + s*crossD)*x (the system matrix of cpp-code/solvers.cpp:144) */
int mvtv_apply_M(mvtv_plan *plan, double s, const double *x, double *out);
/* Replaces: softthresh (cpp-code/solvers.hpp:22) */
int mvtv_softthresh(int64_t n, const double *z, double lam, double *out);
/* Replaces: adapt_step (cpp-code/solvers.hpp:77-82, solvers.cpp:70-88 for MVTV_MODE_CPP: r > 20 s -> rho x20, u x0.05;
 * s > 20 r -> rho x0.1, u x10 ; rcpp solvers.cpp:77-94 for MVTV_MODE_RCPP: factor 10, tau = 2).  r: primal residual
 * (n_r), s: dual residual (n_s), u (n_u, may be 0).  rho_next is NOT truncated to int (the caller does, cpp :126). */
int mvtv_adapt_step(int mode, int64_t n_r, const double *r, int64_t n_s, const double *s, double rho, int64_t n_u,
                    const double *u, double *rho_next, double *u_next);
/* Replaces: nearest1 (cpp-code/utils.hpp:70) on a tensor-product mesh */
int mvtv_nearest(int p, const int64_t *m, const double *axes, int64_t n, const double *data_colmajor,
                 int64_t *vertex_out);

#ifdef __cplusplus
}
#endif
#endif /* MVTV_H */
