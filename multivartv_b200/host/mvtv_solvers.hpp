// mvtv_solvers.hpp -- header-only C++ mirror of the reference's solver interface for the hot path
// (cpp-code/solvers.hpp and rcpp-code/MultivarTV/src/solvers.hpp), implemented on the C ABI of
// include/mvtv.h (libmvtv_b200.so, CUDA sm_100a).  Same function names, argument meaning and error
// behaviour as upstream, so upstream's C++ tests read the same:
//
//   softthresh        cpp-code/solvers.hpp:22
//   admm_update       cpp-code/solvers.hpp:85          (rcpp: solvers.hpp:100, in namespace mvtv::rcpp)
//   mbs_one           cpp-code/solvers.hpp:89          (rcpp: solvers.hpp:104)
//   mbs_predict, mse, mbs_mse   cpp-code/solvers.hpp:93-97
//   adapt_step        cpp-code/solvers.hpp:77-82       (rcpp: solvers.hpp:80-85, in namespace mvtv::rcpp)
//   create_lambdas, mbs_path, test_mse, mbs_fit_optimal, gen_mesh, gen_ftrue, mbs   cpp-code/solvers.hpp:101-129
//   kfold, kfoldinds, rowmean   cpp-code/utils.cpp:406-436 ; rcpp utils.cpp:357-376
//   rcpp::mbs_impl    rcpp-code/MultivarTV/src/solvers.cpp:305-376 (the Rcpp::List as a struct)
//   create_mesh, create_deltas, nearest1   cpp-code/utils.hpp:59-70
//   prod, range, tensor2vector, vector2tensor, dec2binary, fd_binaries   cpp-code/utils.hpp:17-37 (host index maps)
//
// Upstream passes Armadillo types by value; Armadillo is not a dependency here: `mvtv::vec` / `mvtv::mat`
// are minimal column-major containers with the few members the interface needs (n_rows, n_cols, memptr(),
// operator()(i,j), fill()).  Where <armadillo> exists, arma::vec / arma::mat convert implicitly
// (MVTV_WITH_ARMADILLO is set automatically).  There is no CPU fallback: all arithmetic happens on the GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mvtv.h"

#if !defined(MVTV_NO_ARMADILLO) && defined(__has_include)
#if __has_include(<armadillo>)
#include <armadillo>
#define MVTV_WITH_ARMADILLO 1
#endif
#endif

namespace mvtv {

struct vec {
  std::vector<double> mem;
  vec() {}
  explicit vec(size_t n) : mem(n) {}
  vec(std::initializer_list<double> l) : mem(l) {}
  vec(const std::vector<double> &v) : mem(v) {}
#ifdef MVTV_WITH_ARMADILLO
  vec(const arma::vec &a) : mem(a.begin(), a.end()) {}
  operator arma::vec() const { return arma::vec(mem); }
#endif
  size_t size() const { return mem.size(); }
  size_t n_rows_() const { return mem.size(); }
  double *memptr() { return mem.data(); }
  const double *memptr() const { return mem.data(); }
  double &operator[](size_t i) { return mem[i]; }
  double operator[](size_t i) const { return mem[i]; }
  double &operator()(size_t i) { return mem[i]; }
  double operator()(size_t i) const { return mem[i]; }
  void fill(double v) { std::fill(mem.begin(), mem.end(), v); }
};

struct mat {  // column-major, like arma::mat
  size_t n_rows = 0, n_cols = 0;
  std::vector<double> mem;
  mat() {}
  mat(size_t r, size_t c) : n_rows(r), n_cols(c), mem(r * c) {}
#ifdef MVTV_WITH_ARMADILLO
  mat(const arma::mat &a) : n_rows(a.n_rows), n_cols(a.n_cols), mem(a.begin(), a.end()) {}
  mat(const arma::fmat &a) : n_rows(a.n_rows), n_cols(a.n_cols), mem(a.begin(), a.end()) {}
#endif
  double &operator()(size_t i, size_t j) { return mem[i + j * n_rows]; }
  double operator()(size_t i, size_t j) const { return mem[i + j * n_rows]; }
  double *memptr() { return mem.data(); }
  const double *memptr() const { return mem.data(); }
};
typedef mat MAT;  // cpp-code/solvers.hpp:12 (fmat upstream: knots are float-rounded by create_mesh below)

inline void check(int code) {
  if (code == MVTV_OK) return;
  const std::string msg = mvtv_last_error();
  if (code == MVTV_ERR_NOT_CONVERGED) throw std::invalid_argument("Failed to converge!");  // cpp-code/solvers.cpp:123
  if (code == MVTV_ERR_DIM_MISMATCH) throw std::logic_error(msg);                          // arma size mismatch
  throw std::runtime_error("mvtv error " + std::to_string(code) + ": " + msg);
}

// ---- utils.hpp -------------------------------------------------------------------------------------
typedef std::vector<int> VEC;  // cpp-code/utils.hpp:11

inline int prod(int p, const VEC &v) {  // cpp-code/utils.cpp:14-22
  int out = 1;
  for (int i = 0; i < p; ++i) out *= v[(size_t)i];
  return out;
}
inline VEC range(int lo, int hi) {  // cpp-code/utils.cpp:32-38
  VEC v((size_t)(hi - lo + 1));
  for (int i = 0; i < hi - lo + 1; ++i) v[(size_t)i] = lo + i;
  return v;
}
// column-major flattening, axis 0 fastest (cpp-code/utils.cpp:40-52)
inline int tensor2vector(int p, const VEC &multi_ind, const VEC &dims) {
  int v = multi_ind[0];
  for (int i = 1; i < p; ++i) v += multi_ind[(size_t)i] * prod(i, dims);
  return v;
}
// cpp-code/utils.cpp:54-71; upstream divides in float (exact for N <= 2^24), integer ceil here
inline VEC vector2tensor(int p, int vec_ind, const VEC &dims) {
  VEC out((size_t)p);
  int ind2 = vec_ind + 1;
  for (int i = p; i > 0; --i) {
    const int dp = prod(i - 1, dims);
    out[(size_t)(i - 1)] = std::max(1, (ind2 + dp - 1) / dp) - 1;
    ind2 -= out[(size_t)(i - 1)] * dp;
  }
  return out;
}
inline VEC dec2binary(int n, int p) {  // cpp-code/utils.cpp:73-89: p digits, most significant first
  VEC b((size_t)p);
  for (int j = 0; j < p; ++j) b[(size_t)j] = (n >> (p - 1 - j)) & 1;
  return b;
}
inline std::vector<VEC> fd_binaries(int p) {  // cpp-code/utils.cpp:91-101: binaries of 1 .. 2^p - 1
  std::vector<VEC> out;
  for (int i = 1; i < (1 << p); ++i) out.push_back(dec2binary(i, p));
  return out;
}

inline double prodd(const vec &a) {  // cpp-code/utils.cpp:24-30
  double p = 1.0;
  for (size_t i = 0; i < a.size(); ++i) p *= a[i];
  return p;
}

// knots of create_mesh (cpp-code/utils.cpp:271-298): linspace(min+EPS, max+EPS, m_k) rounded to float
// (rcpp variant: min-EPS, EPS=1e-4, double: rcpp utils.cpp:234-254)
inline std::vector<std::vector<double>> mesh_axes(const mat &data, const vec &dims, bool rcpp = false) {
  std::vector<std::vector<double>> axes(data.n_cols);
  for (size_t k = 0; k < data.n_cols; ++k) {
    double lo = std::numeric_limits<double>::infinity(), hi = -lo;
    for (size_t i = 0; i < data.n_rows; ++i) {
      lo = std::min(lo, data(i, k));
      hi = std::max(hi, data(i, k));
    }
    const double eps = rcpp ? 0.0001 : 0.01;
    const double a = rcpp ? lo - eps : lo + eps, b = hi + eps;
    const size_t m = (size_t)dims[k];
    axes[k].resize(m);
    const double delta = m > 1 ? (b - a) / double(m - 1) : 0.0;
    for (size_t j = 0; j + 1 < m; ++j) axes[k][j] = a + double(j) * delta;
    axes[k][m - 1] = b;
    if (!rcpp)
      for (double &v : axes[k]) v = (double)(float)v;
  }
  return axes;
}

inline MAT mesh_from_axes(const std::vector<std::vector<double>> &axes) {
  size_t N = 1;
  for (auto &a : axes) N *= a.size();
  MAT mesh(N, axes.size());
  size_t stride = 1;
  for (size_t k = 0; k < axes.size(); ++k) {
    for (size_t i = 0; i < N; ++i) mesh(i, k) = axes[k][(i / stride) % axes[k].size()];
    stride *= axes[k].size();
  }
  return mesh;
}

inline MAT create_mesh(const mat &data, const vec &dims) { return mesh_from_axes(mesh_axes(data, dims)); }

inline vec create_deltas(const mat &data, const vec &dims, double eps = 0.01) {  // cpp-code/utils.cpp:300-307
  vec d(data.n_cols);
  for (size_t k = 0; k < data.n_cols; ++k) {
    double lo = std::numeric_limits<double>::infinity(), hi = -lo;
    for (size_t i = 0; i < data.n_rows; ++i) {
      lo = std::min(lo, data(i, k));
      hi = std::max(hi, data(i, k));
    }
    d[k] = (hi - lo + 2 * eps) / dims[k];
  }
  return d;
}

// the tensor-product knots behind an N x p mesh matrix (the `MAT mesh` argument of mbs_one)
inline std::vector<double> axes_of_mesh(const MAT &mesh, const vec &m) {
  std::vector<double> axes;
  size_t stride = 1;
  for (size_t k = 0; k < m.size(); ++k) {
    for (size_t j = 0; j < (size_t)m[k]; ++j) axes.push_back(mesh(j * stride, k));
    stride *= (size_t)m[k];
  }
  return axes;
}

inline std::vector<long long> nearest1(const mat &data, const MAT &mesh, const vec &m) {  // cpp-code/utils.cpp:323-330
  std::vector<int64_t> mm(m.size());
  for (size_t k = 0; k < m.size(); ++k) mm[k] = (int64_t)m[k];
  std::vector<double> axes = axes_of_mesh(mesh, m);
  std::vector<int64_t> out(data.n_rows);
  check(mvtv_nearest((int)m.size(), mm.data(), axes.data(), (int64_t)data.n_rows, data.memptr(), out.data()));
  return std::vector<long long>(out.begin(), out.end());
}

// ---- solvers.hpp -----------------------------------------------------------------------------------
inline vec softthresh(vec z, double lam) {  // cpp-code/solvers.hpp:22
  vec out(z.size());
  check(mvtv_softthresh((int64_t)z.size(), z.memptr(), lam, out.memptr()));
  return out;
}

// mbs_cache (cpp-code/solvers.hpp:25-34): the operators O, D, Oty, crossO, crossD live on the device
struct mbs_cache {
  mvtv_plan *plan = nullptr;
  int64_t ntheta = 0, rowsD = 0, n = 0;
  std::vector<double> axes;
  mbs_cache() {}
  mbs_cache(const mbs_cache &) = delete;
  mbs_cache &operator=(const mbs_cache &) = delete;
  ~mbs_cache() {
    if (plan) mvtv_plan_destroy(plan);
  }
};
typedef mbs_cache mbs_one_inits;  // cpp-code/solvers.hpp:36-43: same objects, same owner here

typedef struct mbs_one_object {  // cpp-code/solvers.hpp:45-52 (+ rhohat, uhat of rcpp solvers.hpp:52-61)
  MAT mesh;
  vec theta_hat;
  vec fitted;
  mat data;
  vec y;
  vec m;
  double rhohat = 0.0;
  vec uhat;
  int counter = 0;
} mbs_one_object;

// create_cache_objects (cpp-code/solvers.cpp:31-41): O, D, crossD, crossO, Oty.  `deltas` empty = the
// stand-alone mbs_one path (cpp-code/solvers.cpp:141-145).
inline void create_cache_objects(const mat &data, const vec &y, const MAT &mesh, const vec &meshdims,
                                 mbs_cache &inits, const vec &deltas = vec(), int dtype = MVTV_F64,
                                 int variant = MVTV_VARIANT_REFERENCE) {
  mvtv_plan_desc d{};
  d.struct_size = (int32_t)sizeof(d);
  d.p = (int32_t)meshdims.size();
  for (size_t k = 0; k < meshdims.size() && k < MVTV_MAXP; ++k) d.m[k] = (int64_t)meshdims[k];
  d.dtype = dtype;
  d.variant = variant;
  d.device = -1;
  d.rank = 0;
  d.world = 1;
  d.deltas = deltas.size() ? deltas.memptr() : nullptr;
  if (inits.plan) {
    mvtv_plan_destroy(inits.plan);
    inits.plan = nullptr;
  }
  check(mvtv_plan_create(&inits.plan, &d));
  check(mvtv_plan_info(inits.plan, &inits.ntheta, &inits.rowsD, nullptr, nullptr));
  inits.axes = axes_of_mesh(mesh, meshdims);
  inits.n = (int64_t)data.n_rows;
  check(mvtv_plan_set_points(inits.plan, inits.n, data.memptr(), y.memptr(), inits.axes.data()));
}

// Re-bin a new point set on the operators already held by `inits` (the per-fold create_cache_objects +
// fill_cache of rcpp solvers.cpp:347-348 without rebuilding the plan: D depends only on m and deltas)
inline void cache_set_points(mbs_cache &inits, const mat &data, const vec &y) {
  if (!inits.plan) throw std::logic_error("cache_set_points: empty cache");
  inits.n = (int64_t)data.n_rows;
  check(mvtv_plan_set_points(inits.plan, inits.n, data.memptr(), y.memptr(), inits.axes.data()));
}

inline mvtv_solve_params default_params(int mode, double lambda) {
  mvtv_solve_params p{};
  p.struct_size = (int32_t)sizeof(p);
  p.mode = mode;
  p.lambda = lambda;
  p.rho_init = p.rho_matrix0 = p.tol = std::numeric_limits<double>::quiet_NaN();
  p.precond = MVTV_PRECOND_AUTO;   // like the Python mirror: Jacobi for easy x-updates, the polynomial preconditioner otherwise
  return p;
}

// admm_update (cpp-code/solvers.hpp:85): y is only used for mean(y), which the cache already holds.
inline vec admm_update(const vec & /*y*/, mbs_cache &inits, vec *theta_init, double lambda, int *counter = nullptr) {
  mvtv_solve_params p = default_params(MVTV_MODE_CPP, lambda);
  mvtv_solve_result r{};
  vec theta((size_t)inits.ntheta);
  check(mvtv_solve(inits.plan, &p, theta_init ? theta_init->memptr() : nullptr, nullptr, theta.memptr(), nullptr, &r));
  std::printf("Lambda = %f, Counter = %i \n", lambda, r.counter);  // cpp-code/solvers.cpp:128
  if (counter) *counter = r.counter;
  return theta;
}

// mbs_one (cpp-code/solvers.hpp:89)
inline void mbs_one(const mat &data, const vec &y, const vec &m, mbs_one_object &output, const MAT &mesh,
                    vec *theta_init = NULL, double lambda = 1.0, mbs_cache *cache = NULL) {
  mbs_cache local;
  mbs_cache *inits = cache;
  if (cache == NULL) {  // cpp-code/solvers.cpp:141-145
    create_cache_objects(data, y, mesh, m, local);
    inits = &local;
  }
  mvtv_solve_params p = default_params(MVTV_MODE_CPP, lambda);
  mvtv_solve_result r{};
  output.theta_hat = vec((size_t)inits->ntheta);
  output.fitted = vec((size_t)inits->n);
  check(mvtv_solve(inits->plan, &p, theta_init ? theta_init->memptr() : nullptr, nullptr,
                   output.theta_hat.memptr(), output.fitted.memptr(), &r));
  std::printf("Lambda = %f, Counter = %i \n", lambda, r.counter);
  output.mesh = mesh;  // fill_output_mbs_one, cpp-code/solvers.cpp:64-68
  output.data = data;
  output.y = y;
  output.m = m;
  output.rhohat = r.rho;
  output.counter = r.counter;
}

inline vec mbs_predict(const mbs_one_object &model, const mat &data) {  // cpp-code/solvers.hpp:93
  mbs_cache tmp;
  mvtv_plan_desc d{};
  d.struct_size = (int32_t)sizeof(d);
  d.p = (int32_t)model.m.size();
  for (size_t k = 0; k < model.m.size() && k < MVTV_MAXP; ++k) d.m[k] = (int64_t)model.m[k];
  d.dtype = MVTV_F64;
  d.device = -1;
  d.world = 1;
  check(mvtv_plan_create(&tmp.plan, &d));
  std::vector<double> axes = axes_of_mesh(model.mesh, model.m);
  vec fits(data.n_rows);
  check(mvtv_predict(tmp.plan, (int64_t)data.n_rows, data.memptr(), axes.data(), model.theta_hat.memptr(), fits.memptr()));
  return fits;
}

inline double mse(const vec &fits, const vec &y) {  // cpp-code/solvers.cpp:160-163
  double s = 0.0;
  for (size_t i = 0; i < y.size(); ++i) s += (fits[i] - y[i]) * (fits[i] - y[i]);
  return s / double(y.size());
}
inline double mbs_mse(const mbs_one_object &model, const vec &y) { return mse(model.fitted, y); }

// ---- lambda path and lambda grid (cpp-code/solvers.hpp:54-62,101-109) -------------------------------------
typedef std::vector<mbs_one_object> MBSVEC;
typedef struct mbs_object {  // cpp-code/solvers.hpp:56-62
  mbs_one_object minmse_model;
  MBSVEC models;
  double minmse = 0.0;
  double minmse_lambda = 0.0;
  vec mses;
  vec rhos;  // final rho of every solve of the path (RCPP mode carries it to the next lambda, rcpp solvers.cpp:218)
} mbs_object;

// lam_max_pinv (cpp-code/utils.cpp:399-404) on the cached operators, then the grid of create_lambdas
// (cpp-code/solvers.cpp:179-192): flipud(exp(linspace(log(1e-5*lambda_max), log(lambda_max), n_lambda)))
// (rcpp solvers.cpp:186-200: CGNR lambda_max, grid from 1e-4*lambda_max)
inline vec create_lambdas(int n_lambda, mbs_cache &inits, vec *lambdas = NULL, int mode = MVTV_MODE_CPP) {
  if (lambdas != NULL) return *lambdas;
  double lambda_max = 0.0;
  check(mvtv_lambda_max(inits.plan, mode, &lambda_max, nullptr));
  std::printf("lambda_max = %f ", lambda_max);
  vec out((size_t)n_lambda);
  const double a = std::log(lambda_max * (mode == MVTV_MODE_RCPP ? 0.0001 : 0.00001)), b = std::log(lambda_max);
  for (int i = 0; i < n_lambda; ++i) {
    const double t = (n_lambda == 1 || i == n_lambda - 1) ? b : a + double(i) * ((b - a) / double(n_lambda - 1));
    out[(size_t)(n_lambda - 1 - i)] = std::exp(t);
  }
  return out;
}

// mbs_path (cpp-code/solvers.hpp:109 / solvers.cpp:196-217): warm-started path that stays on the device
inline void mbs_path(const mat &data, const vec &y, const vec &m, const MAT &mesh, int n_lambda, const vec &lambdas,
                     const vec &ftrue, mbs_object &output, mbs_cache &inits, int mode = MVTV_MODE_CPP) {
  mvtv_solve_params p = default_params(mode, lambdas[0]);
  mvtv_solve_result total{};
  const size_t N = (size_t)inits.ntheta, n = (size_t)inits.n;
  std::vector<double> thetas((size_t)n_lambda * N);
  std::vector<int32_t> counters((size_t)n_lambda);
  vec MSEs((size_t)n_lambda);
  output.rhos = vec((size_t)n_lambda);
  int32_t best = 0;
  check(mvtv_solve_path(inits.plan, &p, n_lambda, lambdas.memptr(), ftrue.memptr(), MSEs.memptr(), counters.data(),
                        output.rhos.memptr(), thetas.data(), nullptr, nullptr, &best, &total));
  for (int i = 0; i < n_lambda; ++i) {
    std::printf("Lambda = %f, Counter = %i \n", lambdas[(size_t)i], counters[(size_t)i]);
    mbs_one_object model;
    model.mesh = mesh;
    model.theta_hat = vec(std::vector<double>(thetas.begin() + (size_t)i * N, thetas.begin() + (size_t)(i + 1) * N));
    model.fitted = vec(n);
    check(mvtv_predict(inits.plan, (int64_t)n, data.memptr(), inits.axes.data(), model.theta_hat.memptr(),
                       model.fitted.memptr()));
    model.data = data;
    model.y = y;
    model.m = m;
    model.counter = counters[(size_t)i];
    output.models.push_back(model);
  }
  // fill_output_mbs (cpp-code/solvers.cpp:170-177): first instance of the lowest MSE
  output.minmse_model = output.models[(size_t)best];
  output.minmse = MSEs[(size_t)best];
  output.minmse_lambda = lambdas[(size_t)best];
  output.mses = MSEs;
}

// ---- adapt_step (cpp-code/solvers.hpp:77-82) ---------------------------------------------------------------
typedef struct adaptstep {
  double rho_next = 0.0;
  vec u_next;
} adaptstep;

inline void adapt_step_mode(int mode, const vec &r_current, const vec &s_current, double rho_current, const vec &u_current,
                            adaptstep &object) {
  object.u_next = vec(u_current.size());
  check(mvtv_adapt_step(mode, (int64_t)r_current.size(), r_current.memptr(), (int64_t)s_current.size(), s_current.memptr(),
                        rho_current, (int64_t)u_current.size(), u_current.memptr(), &object.rho_next, object.u_next.memptr()));
}
inline void adapt_step(const vec &r_current, const vec &s_current, double rho_current, const vec &u_current,
                       adaptstep &object) {  // cpp-code/solvers.cpp:70-88
  adapt_step_mode(MVTV_MODE_CPP, r_current, s_current, rho_current, u_current, object);
}

// ---- cross-validation helpers (cpp-code/utils.cpp:406-436 ; rcpp utils.cpp:357-376) -------------------------
inline vec rowmean(const mat &A) {  // cpp-code/utils.cpp:406-413
  vec r(A.n_rows);
  for (size_t i = 0; i < A.n_rows; ++i) {
    double sacc = 0.0;
    for (size_t j = 0; j < A.n_cols; ++j) sacc += A(i, j);
    r[i] = sacc / double(A.n_cols);
  }
  return r;
}

inline mat take_rows(const mat &A, const std::vector<size_t> &ids) {
  mat out(ids.size(), A.n_cols);
  for (size_t j = 0; j < A.n_cols; ++j)
    for (size_t i = 0; i < ids.size(); ++i) out(i, j) = A(ids[i], j);
  return out;
}
inline vec take_rows(const vec &a, const std::vector<size_t> &ids) {
  vec out(ids.size());
  for (size_t i = 0; i < ids.size(); ++i) out[i] = a[ids[i]];
  return out;
}

// arma::shuffle's RNG stream is not reproducible outside Armadillo: a seeded Fisher-Yates on splitmix64 instead
inline uint64_t splitmix64(uint64_t &state) {
  uint64_t z = (state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <typename V>
inline void seeded_shuffle(V &v, uint64_t seed) {
  uint64_t st = seed;
  for (size_t i = v.size(); i > 1; --i) std::swap(v[i - 1], v[(size_t)(splitmix64(st) % (uint64_t)i)]);
}

// kfoldinds (rcpp utils.cpp:367-376): fold label i % k per row, shuffled
inline std::vector<int> kfoldinds(int n, int k, uint64_t seed = 117) {
  std::vector<int> idx((size_t)n);
  for (int i = 0; i < n; ++i) idx[(size_t)i] = i % k;
  seeded_shuffle(idx, seed);
  return idx;
}

// kfold (cpp-code/utils.cpp:417-436): rows shuffled once, fold i tests on the i-th block of n/k rows
typedef struct kfolds {
  std::vector<mat> Xtrain, Xtest;
  std::vector<vec> Ytrain, Ytest;
} kfolds;
inline void kfold(int k, const mat &data, const vec &y, kfolds &struck, uint64_t seed = 117) {
  const size_t n = data.n_rows, ntest = n / (size_t)k;
  std::vector<size_t> perm(n);
  for (size_t i = 0; i < n; ++i) perm[i] = i;
  seeded_shuffle(perm, seed);
  for (int i = 0; i < k; ++i) {
    const size_t first = (size_t)i * ntest, last = (size_t)(i + 1) * ntest;
    std::vector<size_t> te(perm.begin() + (long)first, perm.begin() + (long)last), tr(perm.begin(), perm.begin() + (long)first);
    tr.insert(tr.end(), perm.begin() + (long)last, perm.end());
    struck.Xtest.push_back(take_rows(data, te));
    struck.Ytest.push_back(take_rows(y, te));
    struck.Xtrain.push_back(take_rows(data, tr));
    struck.Ytrain.push_back(take_rows(y, tr));
  }
}

inline MAT gen_mesh(const mat &data, const vec &m, MAT *mesh, bool rcpp = false) {  // cpp-code/solvers.cpp:221-231
  if (mesh != NULL) return *mesh;
  return mesh_from_axes(mesh_axes(data, m, rcpp));
}
inline vec gen_ftrue(const vec &y, vec *ftrue) { return ftrue == NULL ? y : *ftrue; }  // cpp-code/solvers.cpp:235-244

// test_mse (cpp-code/solvers.cpp:264-273): predict the held-out points with every model of the path
inline vec test_mse(const mat &data, const vec &y, const mbs_object &path_object, int n_lambda) {
  vec mses((size_t)n_lambda);
  if (n_lambda == 0) return mses;
  const std::vector<long long> idx = nearest1(data, path_object.models[0].mesh, path_object.models[0].m);
  for (int i = 0; i < n_lambda; ++i) {  // O*theta for a one-hot O is a gather at the nearest vertices
    const vec &th = path_object.models[(size_t)i].theta_hat;
    double sacc = 0.0;
    for (size_t j = 0; j < idx.size(); ++j) sacc += (th[(size_t)idx[j]] - y[j]) * (th[(size_t)idx[j]] - y[j]);
    mses[(size_t)i] = sacc / double(y.size());
  }
  return mses;
}

inline size_t index_min(const vec &v) {  // first minimum, like arma's index_min / find(v - min(v) == 0)[0]
  size_t b = 0;
  for (size_t i = 1; i < v.size(); ++i)
    if (v[i] < v[b]) b = i;
  return b;
}

// mbs_fit_optimal (cpp-code/solvers.cpp:248-260 ; rcpp solvers.cpp:261-274): refit at the lambda with the lowest mean MSE.
// The cached system matrix is whatever the last solve of mbs_path left behind (cpp: crossO + lambdas[last]*crossD,
// warm start from the path's theta; rcpp: cold start, rho = lambdas[0]/5, matrix scalar = the rho the path's
// second-to-last solve ended with).
inline void mbs_fit_optimal(const mat &data, const vec &y, const vec &m, mbs_one_object &best_model, const MAT &mesh,
                            const vec &lambdas, const mat &mse_mat, mbs_cache &cache, const mbs_object &path_object,
                            int mode = MVTV_MODE_CPP) {
  const vec mean_mses = rowmean(mse_mat);
  const size_t best = index_min(mean_mses), nl = lambdas.size();
  mvtv_solve_params p = default_params(mode, lambdas[best]);
  mvtv_solve_result r{};
  best_model.theta_hat = vec((size_t)cache.ntheta);
  best_model.fitted = vec((size_t)cache.n);
  int code;
  if (mode == MVTV_MODE_RCPP) {
    p.rho_init = lambdas[0] / 5.0;
    p.rho_matrix0 = (nl >= 2 && path_object.rhos.size() == nl) ? path_object.rhos[nl - 2] : lambdas[0] / 5.0;
    code = mvtv_solve(cache.plan, &p, nullptr, nullptr, best_model.theta_hat.memptr(), best_model.fitted.memptr(), &r);
    if (code == MVTV_ERR_NOT_CONVERGED) code = MVTV_OK;  // rcpp solvers.cpp:129-132: message + break
  } else {
    p.rho_matrix0 = lambdas[nl - 1];
    code = mvtv_solve(cache.plan, &p, path_object.models[best].theta_hat.memptr(), nullptr, best_model.theta_hat.memptr(),
                      best_model.fitted.memptr(), &r);
  }
  check(code);
  std::printf("Lambda = %f, Counter = %i \n", lambdas[best], r.counter);
  best_model.mesh = mesh;
  best_model.data = data;
  best_model.y = y;
  best_model.m = m;
  best_model.rhohat = r.rho;
  best_model.counter = r.counter;
}

// What rcpp's mbs_impl hands back to R (Rcpp::List, rcpp solvers.cpp:368-373), as a struct
typedef struct mbs_cv_object {
  mbs_one_object best_model;  // data, fitted, m, mesh, theta_hat, y
  vec residuals;              // y - fitted
  MBSVEC models;              // the final path on the full data; models[i].theta_hat / fitted
  vec lambdas, path_mses;     // lambda and mse of every model of the final path
  int lambda_minmse_ind = 0;  // 1-based, like the R list
  vec cv_mses;                // "cv.mses": mean MSE over the folds per lambda
  mat mse_mat;                // n_lambda x folds
} mbs_cv_object;

// The model-selection driver behind mbs (cpp-code/solvers.cpp:277-310) and mbs_impl (rcpp solvers.cpp:305-376).
// Data flow follows rcpp-code, which repairs cpp-code's CV (there every fold re-used the full-data operators and
// path_object.models was never cleared): per-fold operators, a fresh path per fold, final path on the full data.
// `mode` selects the ADMM loop and the mesh / delta / lambda-grid conventions of that sibling.  `foldinds` (length n,
// values 0..folds-1) overrides the seeded shuffle.
inline void mbs_cv(const mat &data, const vec &y, const vec &m, mbs_cv_object &out, MAT *mesh, int n_lambda, vec *ftrue,
                   vec *lambdas, int folds, int mode, bool verbose = true, const std::vector<int> *foldinds = nullptr,
                   uint64_t seed = 117) {
  const bool rc = (mode == MVTV_MODE_RCPP);
  const vec deltas = create_deltas(data, m, rc ? 0.0001 : 0.01);  // inits.deltas (cpp :281 / rcpp :310)
  const MAT MESH = gen_mesh(data, m, mesh, rc);
  if (verbose) std::printf("MBS BEGINS: ntheta = %i \n", (int)prodd(m));
  mbs_cache cache;
  create_cache_objects(data, y, MESH, m, cache, deltas);  // + fill_cache
  const vec LAMBDAS = create_lambdas(n_lambda, cache, lambdas, mode);
  n_lambda = (int)LAMBDAS.size();
  const vec FTRUE = gen_ftrue(y, ftrue);
  mbs_object final_path;
  const int ncol = folds < 1 ? 1 : folds;
  out.mse_mat = mat((size_t)n_lambda, (size_t)ncol);
  if (folds <= 1) {
    mbs_path(data, y, m, MESH, n_lambda, LAMBDAS, FTRUE, final_path, cache, mode);
    for (int i = 0; i < n_lambda; ++i) out.mse_mat((size_t)i, 0) = mse(final_path.models[(size_t)i].fitted, y);  // test_mse on the training data (rcpp :330)
    mbs_fit_optimal(data, y, m, out.best_model, MESH, LAMBDAS, out.mse_mat, cache, final_path, mode);
    out.cv_mses = rowmean(out.mse_mat);
  } else {
    std::vector<int> labels = foldinds ? *foldinds : kfoldinds((int)data.n_rows, folds, seed);
    if (labels.size() != data.n_rows) throw std::invalid_argument("mbs: foldinds must have one label per row of data");
    for (int f = 0; f < folds; ++f) {
      std::vector<size_t> tr, te;
      for (size_t i = 0; i < labels.size(); ++i) (labels[i] == f ? te : tr).push_back(i);
      if (te.empty() || tr.empty()) throw std::invalid_argument("mbs: empty fold");
      const mat train_x = take_rows(data, tr), test_x = take_rows(data, te);
      const vec train_y = take_rows(y, tr), test_y = take_rows(y, te);
      cache_set_points(cache, train_x, train_y);  // per-fold operators (rcpp :347-348)
      mbs_object path_object;
      mbs_path(train_x, train_y, m, MESH, n_lambda, LAMBDAS, train_y, path_object, cache, mode);
      if (verbose) std::printf("Fold = %i Done \n", f + 1);
      const vec col = test_mse(test_x, test_y, path_object, n_lambda);
      for (int i = 0; i < n_lambda; ++i) out.mse_mat((size_t)i, (size_t)f) = col[(size_t)i];
    }
    cache_set_points(cache, data, y);  // final path on the full data (rcpp :355-358)
    mbs_path(data, y, m, MESH, n_lambda, LAMBDAS, y, final_path, cache, mode);
    out.cv_mses = rowmean(out.mse_mat);
    out.best_model = final_path.models[index_min(out.cv_mses)];
  }
  out.lambda_minmse_ind = (int)index_min(out.cv_mses) + 1;
  out.models = final_path.models;
  out.lambdas = LAMBDAS;
  out.path_mses = final_path.mses;
  out.residuals = vec(y.size());
  for (size_t i = 0; i < y.size(); ++i) out.residuals[i] = y[i] - out.best_model.fitted[i];
}

// mbs (cpp-code/solvers.hpp:129): cross-validated fit, best model returned through `output`
inline void mbs(const mat &data, const vec &y, const vec &m, mbs_one_object &output, MAT *mesh = NULL, int n_lambda = 100,
                vec *ftrue = NULL, vec *lambdas = NULL, int folds = 5) {
  mbs_cv_object cv;
  mbs_cv(data, y, m, cv, mesh, n_lambda, ftrue, lambdas, folds, MVTV_MODE_CPP);
  output = cv.best_model;
}

// ---- rcpp-code/MultivarTV/src/solvers.hpp variants ---------------------------------------------------
namespace rcpp {
typedef struct admm_out {  // rcpp solvers.hpp:91-95
  double rho;
  vec theta;
  vec u;
} admm_out;

// admm_update (rcpp solvers.hpp:100).  The first pass uses the cached matrix crossO + rho_init*crossD,
// as mbs_path sets it (rcpp solvers.cpp:213); pass matrix_scalar to override (stand-alone mbs_one: lambda).
inline void admm_update(const vec & /*y*/, mbs_cache &inits, vec &theta_init, double lambda, bool verbose, vec &u_init,
                        double &rho_init, admm_out &out, double matrix_scalar = std::numeric_limits<double>::quiet_NaN(),
                        int *counter = nullptr) {
  mvtv_solve_params p = default_params(MVTV_MODE_RCPP, lambda);
  p.rho_init = rho_init;
  p.rho_matrix0 = (matrix_scalar == matrix_scalar) ? matrix_scalar : rho_init;
  mvtv_solve_result r{};
  out.theta = vec((size_t)inits.ntheta);
  out.u = u_init;
  const int code = mvtv_solve(inits.plan, &p, theta_init.memptr(), out.u.memptr(), out.theta.memptr(), nullptr, &r);
  if (code == MVTV_ERR_NOT_CONVERGED)  // rcpp solvers.cpp:129-132: message + break
    std::printf("ADMM reached max_counter at lambda = %g\n", lambda);
  else
    check(code);
  if (verbose) std::printf("Lambda= %g, Counter = %d\n", lambda, r.counter);
  out.rho = r.rho;
  if (counter) *counter = r.counter;
}

// mbs_one (rcpp solvers.hpp:104)
inline void mbs_one(const mat &data, const vec &y, const vec &m, mbs_one_object &output, const MAT &mesh, vec &u,
                    double &rho, vec &theta_init, double lambda = 1.0, mbs_cache *cache = NULL, bool verbose = true) {
  mbs_cache local;
  mbs_cache *inits = cache;
  double matrix_scalar = std::numeric_limits<double>::quiet_NaN();
  if (cache == NULL) {  // rcpp solvers.cpp:147-151
    create_cache_objects(data, y, mesh, m, local);
    inits = &local;
    matrix_scalar = lambda;
  }
  admm_out out;
  int counter = 0;
  admm_update(y, *inits, theta_init, lambda, verbose, u, rho, out, matrix_scalar, &counter);
  output.mesh = mesh;  // fill_output_mbs_one, rcpp solvers.cpp:71-75
  output.theta_hat = out.theta;
  output.uhat = out.u;
  output.rhohat = out.rho;
  output.fitted = vec((size_t)inits->n);
  check(mvtv_predict(inits->plan, inits->n, data.memptr(), inits->axes.data(), out.theta.memptr(), output.fitted.memptr()));
  output.data = data;
  output.y = y;
  output.m = m;
  output.counter = counter;
}

inline void adapt_step(const vec &r_current, const vec &s_current, double rho_current, const vec &u_current,
                       adaptstep &object) {  // rcpp solvers.cpp:77-94
  adapt_step_mode(MVTV_MODE_RCPP, r_current, s_current, rho_current, u_current, object);
}

// mbs_impl (rcpp solvers.cpp:305-376): what R's mvtv() receives, as a struct
inline mbs_cv_object mbs_impl(const mat &data, const vec &y, const vec &m, MAT *mesh = NULL, int n_lambda = 100,
                              vec *ftrue = NULL, vec *lambdas = NULL, int folds = 5, bool verbose = true,
                              const std::vector<int> *foldinds = nullptr) {
  mbs_cv_object cv;
  mbs_cv(data, y, m, cv, mesh, n_lambda, ftrue, lambdas, folds, MVTV_MODE_RCPP, verbose, foldinds);
  return cv;
}
}  // namespace rcpp

}  // namespace mvtv
