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
  std::printf("utils_test ok\n");
  return 0;
}
