// The flow of upstream's cpp-code/solvers_test.cpp (softthresh, mbs with all defaults, mse on 10000 uniform points
// and a 20x20 mesh, wall time printed) against the C++ mirror, with the assertions upstream leaves out.
// Upstream draws X, y from arma::randu (unseeded); here a seeded splitmix64 stream.
#include <cmath>
#include <ctime>
#include <iostream>

#include "../../multivartv_b200/host/mvtv_solvers.hpp"
using namespace mvtv;

static double randu(uint64_t &st) { return (double)(splitmix64(st) >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char **argv) {
  const size_t n = argc > 1 ? (size_t)std::atol(argv[1]) : 10000;
  const int n_lambda = argc > 2 ? std::atoi(argv[2]) : 100;
  uint64_t st = 117;
  mat X(n, 2);
  vec y(n);
  for (size_t j = 0; j < 2; ++j)
    for (size_t i = 0; i < n; ++i) X(i, j) = randu(st);
  for (size_t i = 0; i < n; ++i) y[i] = randu(st);
  double lam = 0.9;

  vec softy = softthresh(y, lam);
  for (size_t i = 0; i < n; ++i) {  // y >= 0 here: sign(y)*max(|y| - lam, 0)
    const double want = y[i] - lam > 0.0 ? y[i] - lam : 0.0;
    if (softy[i] != want) { std::printf("softthresh wrong at %zu\n", i); return 1; }
  }

  vec m = {20, 20};
  mbs_one_object model_tuned;
  std::clock_t start = std::clock();
  mbs(X, y, m, model_tuned, NULL, n_lambda);
  std::cout << "Time: " << (std::clock() - start) / (double)(CLOCKS_PER_SEC / 1000) << " ms" << std::endl;
  double train_mse = mse(model_tuned.fitted, y);
  std::printf("Tuned model Training MSE = %f \n", train_mse);

  // what upstream only prints: the fit is a 400-vertex piecewise-constant function, fitted = O*theta_hat, and a
  // constant fit (theta = mean(y), the lambda_max end of the path) bounds the training error from above
  if (model_tuned.theta_hat.size() != 400 || model_tuned.fitted.size() != n) { std::printf("bad sizes\n"); return 1; }
  vec again = mbs_predict(model_tuned, X);
  for (size_t i = 0; i < n; ++i)
    if (again[i] != model_tuned.fitted[i]) { std::printf("fitted != O*theta_hat at %zu\n", i); return 1; }
  double meany = 0.0, var = 0.0;
  for (size_t i = 0; i < n; ++i) meany += y[i];
  meany /= (double)n;
  for (size_t i = 0; i < n; ++i) var += (y[i] - meany) * (y[i] - meany);
  var /= (double)n;
  if (!(train_mse <= var * (1.0 + 1e-9)) || !(train_mse > 0.0)) { std::printf("training MSE %g vs var(y) %g\n", train_mse, var); return 1; }
  std::printf("solvers_test ok\n");
  return 0;
}
