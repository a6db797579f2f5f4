// Drop-in check of the C++ mirror (multivartv_b200/host/mvtv_solvers.hpp): written the way upstream's
// cpp-code/solvers_test.cpp uses the interface (softthresh, mbs_one, mse), plus file I/O so the Python parity
// test can compare theta with the oracle.   usage: mbs_one_cli <in.bin> <out.bin>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../multivartv_b200/host/mvtv_solvers.hpp"
using namespace mvtv;

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int64_t n, p, mode;
  double lambda;
  in.read((char *)&n, 8); in.read((char *)&p, 8); in.read((char *)&mode, 8); in.read((char *)&lambda, 8);
  vec m((size_t)p);
  for (int64_t k = 0; k < p; ++k) { int64_t v; in.read((char *)&v, 8); m[k] = (double)v; }
  mat X((size_t)n, (size_t)p);
  vec y((size_t)n);
  in.read((char *)X.memptr(), 8 * n * p);
  in.read((char *)y.memptr(), 8 * n);

  vec softy = softthresh(y, 0.9);                       // cpp-code/solvers_test.cpp:21
  for (size_t i = 0; i < y.size(); ++i) {
    double a = std::fabs(y[i]) - 0.9; a = a > 0 ? a : 0; a = y[i] < 0 ? -a : a;
    if (softy[i] != a) { std::printf("softthresh mismatch at %zu\n", i); return 1; }
  }

  mbs_one_object model;
  try {
    if (mode == 0) {                                    // cpp-code: mbs_one(X, y, m, model, mesh, NULL, lambda)
      MAT mesh = create_mesh(X, m);
      mbs_one(X, y, m, model, mesh, NULL, lambda);
    } else {                                            // rcpp-code: mbs_one(X, y, m, model, mesh, u, rho, theta_init, lambda)
      MAT mesh = mesh_from_axes(mesh_axes(X, m, /*rcpp=*/true));
      double meany = 0; for (size_t i = 0; i < y.size(); ++i) meany += y[i]; meany /= (double)y.size();
      vec theta_init((size_t)prodd(m)); theta_init.fill(meany);
      mbs_cache cache;
      create_cache_objects(X, y, mesh, m, cache);
      vec u((size_t)cache.rowsD); u.fill(0.0);
      double rho = lambda / 5.0;
      rcpp::mbs_one(X, y, m, model, mesh, u, rho, theta_init, lambda, &cache, true);
    }
  } catch (const std::invalid_argument &e) {
    std::printf("caught: %s\n", e.what());
    return 3;
  }
  double train_mse = mse(model.fitted, y);              // cpp-code/solvers_test.cpp:38
  std::printf("Tuned model Training MSE = %f \n", train_mse);
  vec again = mbs_predict(model, X);
  for (size_t i = 0; i < again.size(); ++i) if (again[i] != model.fitted[i]) { std::printf("predict mismatch\n"); return 1; }

  if (mode == 0) {  // mbs_path (cpp-code/solvers.cpp:196-217) on a cache, written the way mbs() calls it
    MAT mesh = create_mesh(X, m);
    mbs_cache cache;
    create_cache_objects(X, y, mesh, m, cache);
    vec lambdas = {4.0, 2.0, lambda};
    mbs_object path;
    mbs_path(X, y, m, mesh, 3, lambdas, y, path, cache);
    if (path.models.size() != 3 || path.mses.size() != 3) { std::printf("mbs_path size mismatch\n"); return 1; }
    double best = path.mses[0];
    for (size_t i = 1; i < 3; ++i) best = std::min(best, path.mses[i]);
    if (path.minmse != best) { std::printf("minmse mismatch\n"); return 1; }
    if (std::fabs(mse(path.models[2].fitted, y) - path.mses[2]) > 1e-12) { std::printf("path mse mismatch\n"); return 1; }
  }
  std::ofstream out(argv[2], std::ios::binary);
  int64_t counter = model.counter, N = (int64_t)model.theta_hat.size();
  out.write((char *)&counter, 8); out.write((char *)&N, 8); out.write((char *)&n, 8);
  out.write((char *)model.theta_hat.memptr(), 8 * N);
  out.write((char *)model.fitted.memptr(), 8 * n);
  return 0;
}
