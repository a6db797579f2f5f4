// C ABI over the REFERENCE's own compiled Rcpp-side solver (rcpp-code/MultivarTV/src/utils.cpp + solvers.cpp of
// brayano/MultivarTV, compiled where they lie against the stand-ins oracle/arma_shim/{armadillo,RcppArmadillo.h}).
// TEST INFRASTRUCTURE ONLY -- see ref_api.cpp.  A separate shared library: the two siblings define the same symbols.
#include <unistd.h>

#include <any>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "solvers.hpp"   // /root/reference/rcpp-code/MultivarTV/src/solvers.hpp

using namespace arma;

void create_cache_objects(mat data, vec y, MAT mesh, vec meshdims, mbs_one_inits &inits);
arma::vec create_lambdas(int n_lambda, mbs_one_inits inits, Rcpp::Nullable<arma::vec> lambdas, bool verbose);
void mbs_path(mat data, vec y, vec m, MAT mesh, int n_lambda, vec lambdas, vec ftrue, mbs_object &output, mbs_one_inits inits,
              mbs_cache *cache, bool verbose);
void fill_cache(mbs_cache *&cache, mbs_one_inits inits);
double lam_max_pinv(sp_mat a, vec Oty);
Rcpp::List mbs_impl(const arma::mat data, const arma::vec y, arma::vec m, Rcpp::Nullable<arma::mat> mesh, int n_lambda,
                    Rcpp::Nullable<arma::vec> ftrue, Rcpp::Nullable<arma::vec> lambdas, int folds, bool verbose);

static thread_local std::string g_err;

static mat to_mat(const double *a, long long r, long long c) {
  mat m((uword)r, (uword)c);
  std::memcpy(m.memptr(), a, sizeof(double) * (size_t)(r * c));
  return m;
}
static vec to_vec(const double *a, long long n) {
  vec v((uword)n);
  if (n) std::memcpy(v.memptr(), a, sizeof(double) * (size_t)n);
  return v;
}

struct StdoutCapture {   // "Lambda= <l>, Counter = <c>" on Rcpp::Rcout is the only place the iteration count appears
  int saved = -1;
  FILE *tmp = nullptr;
  StdoutCapture() {
    std::cout.flush();
    fflush(stdout);
    tmp = tmpfile();
    saved = dup(fileno(stdout));
    dup2(fileno(tmp), fileno(stdout));
  }
  std::string finish() {
    std::cout.flush();
    fflush(stdout);
    dup2(saved, fileno(stdout));
    close(saved);
    saved = -1;
    std::string out;
    rewind(tmp);
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), tmp)) > 0) out.append(buf, n);
    fclose(tmp);
    tmp = nullptr;
    return out;
  }
  ~StdoutCapture() {
    if (saved >= 0) finish();
  }
};
static int parse_counters(const std::string &s, int *out, int cap) {
  int n = 0;
  size_t pos = 0;
  const std::string key = "Counter = ";
  while ((pos = s.find(key, pos)) != std::string::npos) {
    pos += key.size();
    if (n < cap) out[n] = atoi(s.c_str() + pos);
    ++n;
  }
  return n;
}

#define REF_GUARD(...)                       \
  try {                                      \
    __VA_ARGS__;                             \
    return 0;                                \
  } catch (const std::invalid_argument &e) { \
    g_err = e.what();                        \
    return 3;                                \
  } catch (const std::logic_error &e) {      \
    g_err = e.what();                        \
    return 4;                                \
  } catch (const std::exception &e) {        \
    g_err = e.what();                        \
    return 1;                                \
  }

extern "C" {

const char *rref_last_error(void) { return g_err.c_str(); }

// create_mesh / create_deltas of the Rcpp side (utils.cpp:234-263: min-EPS .. max+EPS, EPS = 1e-4, double)
int rref_create_mesh(long long n, int p, const double *data, const double *dims, double *mesh_out) {
  REF_GUARD({
    MAT mesh = create_mesh(to_mat(data, n, p), to_vec(dims, p));
    for (size_t i = 0; i < mesh.mem.size(); ++i) mesh_out[i] = (double)mesh.mem[i];
  })
}
int rref_create_deltas(long long n, int p, const double *data, const double *dims, double *out) {
  REF_GUARD({
    vec d = create_deltas(to_mat(data, n, p), to_vec(dims, p));
    for (int k = 0; k < p; ++k) out[k] = d[k];
  })
}
int rref_adapt_step(long long nr, const double *r, long long ns, const double *s, double rho, long long nu, const double *u,
                    double *rho_next, double *u_next) {
  REF_GUARD({
    adaptstep obj;
    adapt_step(to_vec(r, nr), to_vec(s, ns), rho, to_vec(u, nu), obj);
    *rho_next = obj.rho_next;
    std::memcpy(u_next, obj.u_next.memptr(), sizeof(double) * (size_t)nu);
  })
}
int rref_rows_of_D(int p, const double *dims, long long *rows) {
  REF_GUARD({ *rows = (long long)create_D(p, to_vec(dims, p), vec()).n_rows; })
}
// mbs_one (rcpp solvers.cpp:140-159), stand-alone (cache == NULL, matrix crossO + lambda*crossD until the loop rebuilds it):
// theta_init (N), u (R, in/out), rho (in/out) as upstream passes them by reference.
int rref_mbs_one(long long n, int p, const double *data, const double *y, const double *m, const double *theta_init, double *u_inout,
                 double *rho_inout, double lambda, double *theta_out, double *fitted_out, int *counter) {
  StdoutCapture cap;
  int rc = [&]() -> int {
    REF_GUARD({
      mat X = to_mat(data, n, p);
      vec Y = to_vec(y, n), M = to_vec(m, p);
      MAT mesh = create_mesh(X, M);
      const long long N = (long long)prodd(M);
      const long long R = (long long)create_D(p, M, vec()).n_rows;
      vec th0 = to_vec(theta_init, N), u = to_vec(u_inout, R);
      double rho = *rho_inout;
      mbs_one_object out;
      mbs_one(X, Y, M, out, mesh, u, rho, th0, lambda, NULL, true);
      std::memcpy(theta_out, out.theta_hat.memptr(), sizeof(double) * (size_t)N);
      std::memcpy(fitted_out, out.fitted.memptr(), sizeof(double) * (size_t)n);
      std::memcpy(u_inout, out.uhat.memptr(), sizeof(double) * (size_t)R);
      *rho_inout = out.rhohat;
    })
  }();
  int c = 0;
  parse_counters(cap.finish(), &c, 1);
  if (counter) *counter = c;
  return rc;
}
// mbs_impl's operator set-up (rcpp solvers.cpp:307-319) + create_lambdas (:186-200) + mbs_path (:204-222) on the full data
int rref_mbs_path(long long n, int p, const double *data, const double *y, const double *m, int n_lambda, const double *lambdas_in,
                  double *lambdas_out, double *thetas_out, double *mses_out, double *rhos_out, int *counters_out, double *lambda_max_out) {
  StdoutCapture cap;
  int rc = [&]() -> int {
    REF_GUARD({
      mat X = to_mat(data, n, p);
      vec Y = to_vec(y, n), M = to_vec(m, p);
      mbs_one_inits inits;
      inits.ntheta = prodd(M);
      inits.deltas = create_deltas(X, M);
      MAT MESH = create_mesh(X, M);
      mbs_cache *cache = new mbs_cache();
      cache->ntheta = inits.ntheta;
      create_cache_objects(X, Y, MESH, M, inits);
      fill_cache(cache, inits);
      Rcpp::Nullable<arma::vec> L;
      if (lambdas_in) L = Rcpp::Nullable<arma::vec>(to_vec(lambdas_in, n_lambda));
      vec LAMBDAS = create_lambdas(n_lambda, inits, L, true);
      *lambda_max_out = lambdas_in ? NAN : LAMBDAS[0];
      mbs_object path;
      mbs_path(X, Y, M, MESH, n_lambda, LAMBDAS, Y, path, inits, cache, true);
      const size_t N = (size_t)inits.ntheta;
      for (int i = 0; i < n_lambda; ++i) {
        lambdas_out[i] = LAMBDAS[i];
        mses_out[i] = path.mses[i];
        rhos_out[i] = path.models[i].rhohat;
        std::memcpy(thetas_out + (size_t)i * N, path.models[i].theta_hat.memptr(), sizeof(double) * N);
      }
      delete cache;
    })
  }();
  parse_counters(cap.finish(), counters_out, n_lambda);
  return rc;
}
int rref_lambda_max(long long n, int p, const double *data, const double *y, const double *m, double *out) {
  StdoutCapture cap;
  REF_GUARD({
    mat X = to_mat(data, n, p);
    vec Y = to_vec(y, n), M = to_vec(m, p);
    mbs_one_inits inits;
    inits.ntheta = prodd(M);
    inits.deltas = create_deltas(X, M);
    create_cache_objects(X, Y, create_mesh(X, M), M, inits);
    *out = lam_max_pinv(inits.D, inits.Oty);
  })
}
// mbs_impl (rcpp solvers.cpp:305-376) with folds = 1 (no shuffle involved): the list R's mvtv() receives
int rref_mbs_impl_folds1(long long n, int p, const double *data, const double *y, const double *m, int n_lambda, const double *lambdas,
                         double *theta_out, double *fitted_out, double *cv_mses_out, int *best_index_1based) {
  StdoutCapture cap;
  REF_GUARD({
    Rcpp::List res = mbs_impl(to_mat(data, n, p), to_vec(y, n), to_vec(m, p), R_NilValue, n_lambda, R_NilValue,
                              Rcpp::Nullable<arma::vec>(to_vec(lambdas, n_lambda)), 1, true);
    const vec th = std::any_cast<vec>(res["theta_hat"]), fit = std::any_cast<vec>(res["fitted"]);
    std::memcpy(theta_out, th.memptr(), sizeof(double) * (size_t)th.n_elem);
    std::memcpy(fitted_out, fit.memptr(), sizeof(double) * (size_t)fit.n_elem);
    const vec cv = std::any_cast<vec>(res["cv.mses"]);
    for (int i = 0; i < n_lambda; ++i) cv_mses_out[i] = cv[i];
    *best_index_1based = (int)std::any_cast<uword>(res["lambda_minmse_ind"]);
  })
}

}  // extern "C"
