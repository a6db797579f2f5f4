// The known answers upstream's cpp-code/utils_test.cpp only prints ("0 is correct", "27-1 is correct", {0,0,0},
// {2,2,2}, the 7 binaries for p = 3), asserted against the C++ mirror.  Index maps are host bookkeeping: this
// program needs no GPU.
#include <cstdio>

#include "../../multivartv_b200/host/mvtv_solvers.hpp"
using namespace mvtv;

#define EXPECT(cond)                                                   \
  do {                                                                 \
    if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } \
  } while (0)

int main() {
  const int P = 3;
  const VEC dims = {3, 3, 3}, myind = {0, 0, 0}, myind2 = {2, 2, 2};
  EXPECT(tensor2vector(P, myind, dims) == 0);                 // utils_test.cpp:64 "0 is correct"
  EXPECT(tensor2vector(P, myind2, dims) == 26);               // :65 "27-1 is correct"
  EXPECT(vector2tensor(P, 0, dims) == (VEC{0, 0, 0}));        // :67-71
  EXPECT(vector2tensor(P, 26, dims) == (VEC{2, 2, 2}));
  EXPECT(prod(P, myind2) == 8);                               // :54
  EXPECT(range(1, 4) == (VEC{1, 2, 3, 4}));                   // :57-61
  const VEC altdims = {3, 2, 3};                              // :73-82
  EXPECT(vector2tensor(P, 0, altdims) == (VEC{0, 0, 0}));
  EXPECT(vector2tensor(P, 1, altdims) == (VEC{1, 0, 0}));
  EXPECT(vector2tensor(P, 2, altdims) == (VEC{2, 0, 0}));
  for (int v = 0; v < 18; ++v) EXPECT(tensor2vector(P, vector2tensor(P, v, altdims), altdims) == v);
  const char *want[7] = {"001", "010", "011", "100", "101", "110", "111"};   // :84-97
  const std::vector<VEC> bins = fd_binaries(P);
  EXPECT(bins.size() == 7);
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < P; ++j) {
      EXPECT(bins[(size_t)i][(size_t)j] == want[i][j] - '0');
      EXPECT(dec2binary(i + 1, P)[(size_t)j] == want[i][j] - '0');
    }
  // host-side mesh helpers (cpp-code/utils.cpp:271-307): 6 knots on [0.01 + EPS, 0.99 + EPS], float-rounded
  mat data(10, 1);
  for (size_t i = 0; i < 10; ++i) data(i, 0) = 0.01 + 0.98 * (double)i / 9.0;
  const vec m1 = {6};
  const MAT mesh = create_mesh(data, m1);
  EXPECT(mesh.n_rows == 6 && mesh.n_cols == 1);
  EXPECT(mesh(0, 0) == (double)(float)0.02 && mesh(5, 0) == (double)(float)1.0);
  const vec d = create_deltas(data, m1);
  EXPECT(std::fabs(d[0] - (0.98 + 0.02) / 6.0) < 1e-15);
  // cross-validation helpers (cpp-code/utils.cpp:406-436 ; rcpp utils.cpp:367-376): host-only
  {
    const int n = 23, k = 5;
    const std::vector<int> lab = kfoldinds(n, k, 117), lab2 = kfoldinds(n, k, 117), lab3 = kfoldinds(n, k, 118);
    EXPECT(lab.size() == (size_t)n && lab == lab2 && lab != lab3);          // seeded: reproducible, seed-dependent
    int cnt[5] = {0, 0, 0, 0, 0};
    for (int v : lab) { EXPECT(v >= 0 && v < k); cnt[v]++; }
    for (int f = 0; f < k; ++f) EXPECT(cnt[f] == n / k + (f < n % k ? 1 : 0));   // labels are a permutation of i % k
    mat X((size_t)n, 2);
    vec yv((size_t)n);
    for (size_t i = 0; i < (size_t)n; ++i) { X(i, 0) = (double)i; X(i, 1) = 100.0 + (double)i; yv[i] = 1000.0 + (double)i; }
    kfolds folds;
    kfold(k, X, yv, folds, 7);
    EXPECT(folds.Xtest.size() == (size_t)k && folds.Ytrain.size() == (size_t)k);
    std::vector<int> seen((size_t)n, 0);
    for (int f = 0; f < k; ++f) {
      EXPECT(folds.Xtest[(size_t)f].n_rows == (size_t)(n / k) && folds.Xtrain[(size_t)f].n_rows == (size_t)(n - n / k));
      std::vector<int> in_fold((size_t)n, 0);
      for (size_t i = 0; i < folds.Xtest[(size_t)f].n_rows; ++i) {
        const int row = (int)folds.Xtest[(size_t)f](i, 0);
        EXPECT(folds.Xtest[(size_t)f](i, 1) == 100.0 + row && folds.Ytest[(size_t)f][i] == 1000.0 + row);   // rows stay together
        seen[(size_t)row]++; in_fold[(size_t)row]++;
      }
      for (size_t i = 0; i < folds.Xtrain[(size_t)f].n_rows; ++i) {
        const int row = (int)folds.Xtrain[(size_t)f](i, 0);
        EXPECT(folds.Ytrain[(size_t)f][i] == 1000.0 + row);
        in_fold[(size_t)row]++;
      }
      for (int v : in_fold) EXPECT(v == 1);                                   // train and test partition the rows
    }
    int tested = 0;
    for (int v : seen) { EXPECT(v <= 1); tested += v; }
    EXPECT(tested == k * (n / k));                                            // test blocks are disjoint (utils.cpp:425-428)
    mat A(2, 3);
    A(0, 0) = 1; A(0, 1) = 2; A(0, 2) = 6; A(1, 0) = -1; A(1, 1) = 0; A(1, 2) = 1;
    const vec rm = rowmean(A);
    EXPECT(rm[0] == 3.0 && rm[1] == 0.0);
    const vec v3 = {2.0, 1.0, 1.0, 5.0};
    EXPECT(index_min(v3) == 1);                                               // first minimum (find(v - min(v) == 0)[0])
    vec ftrue = {9.0};
    EXPECT(gen_ftrue(v3, NULL).size() == 4 && gen_ftrue(v3, &ftrue)[0] == 9.0);
  }
  std::printf("utils_test ok\n");
  return 0;
}
