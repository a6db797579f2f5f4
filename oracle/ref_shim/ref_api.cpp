// C ABI over the REFERENCE's own compiled C++ solver (cpp-code/utils.cpp + cpp-code/solvers.cpp of brayano/MultivarTV,
// compiled where they lie under /root/reference against the Armadillo stand-in oracle/arma_shim/armadillo).
// TEST INFRASTRUCTURE ONLY: tests/ and bench.py's reference arm call this to pin the restated oracles and to generate
// golden vectors; the product never links it.  Every function below only marshals plain arrays into the reference's
// types and calls the reference function named in its comment.
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "solvers.hpp"   // /root/reference/cpp-code/solvers.hpp (includes utils.hpp)

using namespace arma;

// functions defined in the reference's .cpp files but missing from (or stale in) its headers
void create_cache_objects(mat data, vec y, MAT mesh, vec meshdims, mbs_one_inits &inits);
vec create_lambdas(int n_lambda, mbs_one_inits inits, vec *lambdas);
void mbs_path(mat data, vec y, vec m, MAT mesh, int n_lambda, vec lambdas, vec ftrue, mbs_object &output, mbs_one_inits inits, mbs_cache *cache);

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

// stdout of the reference ("Lambda = %f, Counter = %i") is the only place its iteration count appears: capture it
struct StdoutCapture {
  int saved = -1;
  FILE *tmp = nullptr;
  StdoutCapture() {
    fflush(stdout);
    tmp = tmpfile();
    saved = dup(fileno(stdout));
    dup2(fileno(tmp), fileno(stdout));
  }
  std::string finish() {
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

const char *ref_last_error(void) { return g_err.c_str(); }

// tensor2vector / vector2tensor (cpp-code/utils.cpp:40-71)
int ref_tensor2vector(int p, const int *multi_ind, const int *dims) {
  return tensor2vector(p, VEC(multi_ind, multi_ind + p), VEC(dims, dims + p));
}
void ref_vector2tensor(int p, int vec_ind, const int *dims, int *out) {
  VEC r = vector2tensor(p, vec_ind, VEC(dims, dims + p));
  for (int k = 0; k < p; ++k) out[k] = r[k];
}
// fd_binaries (cpp-code/utils.cpp:91-101): (2^p - 1) x p, row-major out
void ref_fd_binaries(int p, int *out) {
  umat b = fd_binaries(p);
  for (uword i = 0; i < b.n_rows; ++i)
    for (uword j = 0; j < b.n_cols; ++j) out[i * b.n_cols + j] = (int)b(i, j);
}
// create_D (cpp-code/utils.cpp:245-269): dense R x N, row-major out; deltas may be NULL (empty vec, the stand-alone path)
int ref_create_D_rows(int p, const double *dims, const double *deltas, long long *rows, long long *cols) {
  REF_GUARD({
    sp_mat D = create_D(p, to_vec(dims, p), deltas ? to_vec(deltas, p) : vec());
    *rows = (long long)D.n_rows;
    *cols = (long long)D.n_cols;
  })
}
int ref_create_D_dense(int p, const double *dims, const double *deltas, double *out) {
  REF_GUARD({
    sp_mat D = create_D(p, to_vec(dims, p), deltas ? to_vec(deltas, p) : vec());
    std::memset(out, 0, sizeof(double) * (size_t)(D.n_rows * D.n_cols));
    for (uword j = 0; j < D.n_cols; ++j)
      for (uword k = D.col_ptrs[j]; k < D.col_ptrs[j + 1]; ++k) out[D.row_indices[k] * D.n_cols + j] = D.values[k];
  })
}
// create_mesh / create_deltas (cpp-code/utils.cpp:271-307); mesh out N x p column-major (float values widened)
int ref_create_mesh(long long n, int p, const double *data, const double *dims, double *mesh_out) {
  REF_GUARD({
    MAT mesh = create_mesh(to_mat(data, n, p), to_vec(dims, p));
    for (size_t i = 0; i < mesh.mem.size(); ++i) mesh_out[i] = (double)mesh.mem[i];
  })
}
int ref_create_deltas(long long n, int p, const double *data, const double *dims, double *out) {
  REF_GUARD({
    vec d = create_deltas(to_mat(data, n, p), to_vec(dims, p));
    for (int k = 0; k < p; ++k) out[k] = d[k];
  })
}
// nearest1 (cpp-code/utils.cpp:323-330) against an N x p mesh (column-major doubles, narrowed to the reference's fmat)
int ref_nearest1(long long n, int p, const double *data, long long N, const double *mesh, long long *out) {
  REF_GUARD({
    MAT M((uword)N, (uword)p);
    for (size_t i = 0; i < M.mem.size(); ++i) M.mem[i] = (float)mesh[i];
    uvec idx = nearest1(to_mat(data, n, p), M);
    for (long long i = 0; i < n; ++i) out[i] = (long long)idx[i];
  })
}
// softthresh (cpp-code/solvers.cpp:24-29)
int ref_softthresh(long long n, const double *z, double lam, double *out) {
  REF_GUARD({
    vec r = softthresh(to_vec(z, n), lam);
    std::memcpy(out, r.memptr(), sizeof(double) * (size_t)n);
  })
}
// adapt_step (cpp-code/solvers.cpp:70-88)
int ref_adapt_step(long long nr, const double *r, long long ns, const double *s, double rho, long long nu, const double *u,
                   double *rho_next, double *u_next) {
  REF_GUARD({
    adaptstep obj;
    adapt_step(to_vec(r, nr), to_vec(s, ns), rho, to_vec(u, nu), obj);
    *rho_next = obj.rho_next;
    std::memcpy(u_next, obj.u_next.memptr(), sizeof(double) * (size_t)nu);
  })
}
// mbs_one (cpp-code/solvers.cpp:134-152), stand-alone (cache == NULL): mesh = create_mesh(data, m) unless given;
// theta_init may be NULL.  counter = the "Counter" the reference prints.
int ref_mbs_one(long long n, int p, const double *data, const double *y, const double *m, const double *mesh_in /*N x p or NULL*/,
                const double *theta_init, double lambda, double *theta_out, double *fitted_out, int *counter) {
  StdoutCapture cap;
  int rc = [&]() -> int {
    REF_GUARD({
      mat X = to_mat(data, n, p);
      vec Y = to_vec(y, n), M = to_vec(m, p);
      MAT mesh;
      if (mesh_in) {
        const uword N = (uword)prodd(M);
        mesh.set_size(N, (uword)p);
        for (size_t i = 0; i < mesh.mem.size(); ++i) mesh.mem[i] = (float)mesh_in[i];
      } else {
        mesh = create_mesh(X, M);
      }
      vec th0;
      if (theta_init) th0 = to_vec(theta_init, (long long)prodd(M));
      mbs_one_object out;
      mbs_one(X, Y, M, out, mesh, theta_init ? &th0 : NULL, lambda, NULL);
      std::memcpy(theta_out, out.theta_hat.memptr(), sizeof(double) * (size_t)out.theta_hat.n_elem);
      std::memcpy(fitted_out, out.fitted.memptr(), sizeof(double) * (size_t)out.fitted.n_elem);
    })
  }();
  const std::string log = cap.finish();
  int c = 0;
  parse_counters(log, &c, 1);
  if (counter) *counter = c;
  return rc;
}
// The operator set-up of mbs() (cpp-code/solvers.cpp:279-287: deltas, mesh, create_cache_objects, fill_cache), then
// create_lambdas / lam_max_pinv (:179-192, utils.cpp:354-404) and mbs_path (:196-217) on the full data.
// lambdas_in == NULL -> the reference's own grid of n_lambda values.  thetas_out: n_lambda x N (row-major), counters_out,
// mses_out: n_lambda, lambdas_out: n_lambda, lambda_max_out (NaN when the grid was given).
int ref_mbs_path(long long n, int p, const double *data, const double *y, const double *m, int n_lambda, const double *lambdas_in,
                 const double *ftrue, double *lambdas_out, double *thetas_out, double *mses_out, int *counters_out,
                 double *lambda_max_out) {
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
      vec L;
      if (lambdas_in) L = to_vec(lambdas_in, n_lambda);
      vec LAMBDAS = create_lambdas(n_lambda, inits, lambdas_in ? &L : NULL);
      *lambda_max_out = lambdas_in ? NAN : LAMBDAS[0];
      mbs_object path;
      mbs_path(X, Y, M, MESH, n_lambda, LAMBDAS, ftrue ? to_vec(ftrue, n) : Y, path, inits, cache);
      const size_t N = (size_t)inits.ntheta;
      for (int i = 0; i < n_lambda; ++i) {
        lambdas_out[i] = LAMBDAS[i];
        mses_out[i] = path.mses[i];
        std::memcpy(thetas_out + (size_t)i * N, path.models[i].theta_hat.memptr(), sizeof(double) * N);
      }
      delete cache;
    })
  }();
  const std::string log = cap.finish();
  parse_counters(log, counters_out, n_lambda);
  return rc;
}
// lam_max_pinv (cpp-code/utils.cpp:399-404) on the operators of mbs() (deltas set) for the given data
int ref_lambda_max(long long n, int p, const double *data, const double *y, const double *m, double *out) {
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

}  // extern "C"
