// TEST INFRASTRUCTURE ONLY: cub::DeviceRadixSort::SortPairs as a stable host sort, for the CPU emulation of csrc/setup.cu.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "../../cuda_runtime.h"
namespace cub {
struct DeviceRadixSort {
  template <typename K, typename V>
  static cudaError_t SortPairs(void *temp, size_t &temp_bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int = 0, int = 32, cudaStream_t = nullptr) {
    if (!temp) { temp_bytes = 16; return cudaSuccess; }
    std::vector<int> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return kin[a] < kin[b]; });
    for (int i = 0; i < n; ++i) { kout[i] = kin[idx[(size_t)i]]; vout[i] = vin[idx[(size_t)i]]; }
    return cudaSuccess;
  }
};
}  // namespace cub
