// Cross-validated driver of the C++ mirror (mbs / rcpp::mbs_impl, multivartv_b200/host/mvtv_solvers.hpp) with file
// I/O, so the Python parity test can run the same folds through multivartv_b200.mbs and compare.
//   usage: mbs_cli <in.bin> <out.bin>
//   in : int64 n, p, mode (0 cpp / 1 rcpp), folds, n_lambda, has_lambdas ; int64 m[p] ; double X[n*p] (column-major) ;
//        double y[n] ; double lambdas[n_lambda] if has_lambdas ; int64 foldinds[n] if folds > 1
//   out: int64 n_lambda, N, n, lambda_minmse_ind ; double lambdas[nl], cv_mses[nl], mse_mat[nl*max(folds,1)] (column-major),
//        theta_hat[N], fitted[n]
#include <cstdint>
#include <cstdio>
#include <fstream>

#include "../../multivartv_b200/host/mvtv_solvers.hpp"
using namespace mvtv;

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int64_t hdr[6];
  in.read((char *)hdr, sizeof(hdr));
  const int64_t n = hdr[0], p = hdr[1], mode = hdr[2], folds = hdr[3], n_lambda = hdr[4], has_lambdas = hdr[5];
  vec m((size_t)p);
  for (int64_t k = 0; k < p; ++k) { int64_t v; in.read((char *)&v, 8); m[(size_t)k] = (double)v; }
  mat X((size_t)n, (size_t)p);
  vec y((size_t)n);
  in.read((char *)X.memptr(), 8 * n * p);
  in.read((char *)y.memptr(), 8 * n);
  vec lambdas((size_t)(has_lambdas ? n_lambda : 0));
  if (has_lambdas) in.read((char *)lambdas.memptr(), 8 * n_lambda);
  std::vector<int> foldinds;
  if (folds > 1) {
    std::vector<int64_t> f((size_t)n);
    in.read((char *)f.data(), 8 * n);
    foldinds.assign(f.begin(), f.end());
  }
  if (!in) { std::printf("short input file\n"); return 2; }

  mbs_cv_object cv;
  try {
    mbs_cv(X, y, m, cv, NULL, (int)n_lambda, NULL, has_lambdas ? &lambdas : NULL, (int)folds,
           mode == 0 ? MVTV_MODE_CPP : MVTV_MODE_RCPP, true, folds > 1 ? &foldinds : nullptr);
  } catch (const std::invalid_argument &e) {
    std::printf("caught: %s\n", e.what());
    return 3;
  }
  // the reference-shaped entry point returns the same best model (cpp-code/solvers.hpp:129); only checked when the
  // folds are the library's own seeded ones, i.e. reproducible without the injected labels
  if (mode == 0 && folds <= 1) {
    mbs_one_object again;
    mbs(X, y, m, again, NULL, (int)n_lambda, NULL, has_lambdas ? &lambdas : NULL, (int)folds);
    for (size_t i = 0; i < again.theta_hat.size(); ++i)
      if (again.theta_hat[i] != cv.best_model.theta_hat[i]) { std::printf("mbs() differs from mbs_cv()\n"); return 1; }
  }
  std::ofstream out(argv[2], std::ios::binary);
  const int64_t nl = (int64_t)cv.lambdas.size(), N = (int64_t)cv.best_model.theta_hat.size(), best = cv.lambda_minmse_ind;
  out.write((const char *)&nl, 8); out.write((const char *)&N, 8); out.write((const char *)&n, 8); out.write((const char *)&best, 8);
  out.write((const char *)cv.lambdas.memptr(), 8 * nl);
  out.write((const char *)cv.cv_mses.memptr(), 8 * nl);
  out.write((const char *)cv.mse_mat.memptr(), 8 * (int64_t)cv.mse_mat.mem.size());
  out.write((const char *)cv.best_model.theta_hat.memptr(), 8 * N);
  out.write((const char *)cv.best_model.fitted.memptr(), 8 * n);
  std::printf("best lambda index (1-based) = %d, cv mse = %g\n", cv.lambda_minmse_ind, cv.cv_mses[(size_t)(best - 1)]);
  return 0;
}
