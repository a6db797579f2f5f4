// Setup / teardown kernels: the interpolation operator A = O (nearest mesh vertex) and its transpose.
//   nearest1 / nearest_interp_matrix  cpp-code/utils.cpp:311-352  -> k_bin (separable per-axis nearest knot)
//   Oty = Ot*y, crossO = Ot*O         cpp-code/solvers.cpp:37,40  -> radix sort by vertex + k_segment_reduce
//   fitted = O*theta                  cpp-code/solvers.cpp:66     -> k_gather
// No atomics: points are sorted by vertex (stable, so every vertex's points stay in increasing point
// order = the order arma's sparse product accumulates in) and each segment is reduced by one thread.
#include <cub/device/device_radix_sort.cuh>

#include "setup.h"

namespace mvtv {

// nearest knot of a sorted axis by the reference's own metric (x-knot)^2, ties -> lower index
__device__ __forceinline__ long long nearest_knot(const double *__restrict__ ax, long long m, double x) {
  long long lo = 0, hi = m;  // first j with ax[j] >= x
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (ax[mid] < x) lo = mid + 1; else hi = mid;
  }
  long long best = lo < m ? lo : m - 1;
  double bd = (x - ax[best]) * (x - ax[best]);
  for (long long j = best - 1; j >= 0 && j >= best - 2; --j) {
    const double d = (x - ax[j]) * (x - ax[j]);
    if (d <= bd) { bd = d; best = j; }
  }
  return best;
}

// vid[i] = global vertex id of point i; key[i] = local slab vertex (or 0xFFFFFFFF outside the slab)
__global__ void k_bin(int p, DimTab dt, long long n, const double *__restrict__ data, long long ld_point,
                      long long ld_axis, const double *__restrict__ axes, long long *__restrict__ vid,
                      unsigned *__restrict__ key, unsigned *__restrict__ val) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long v = 0, stride = 1, zlast = 0;
    const double *ax = axes;
    for (int a = 0; a < p; ++a) {
      const long long j = nearest_knot(ax, dt.m[a], data[i * ld_point + (long long)a * ld_axis]);
      v += j * stride;
      stride *= dt.m[a];
      ax += dt.m[a];
      zlast = j;
    }
    if (p < dt.P) zlast = 0;  // 1-D meshes are stored as (m, 1)
    if (vid) vid[i] = v;
    if (key) {
      const long long zl = zlast - dt.z0;
      const bool mine = zl >= 0 && zl < dt.nz;
      key[i] = mine ? (unsigned)(v - dt.z0 * dt.plane) : 0xFFFFFFFFu;
      val[i] = (unsigned)i;
    }
  }
}

// one thread per segment head of the sorted keys: sequential (deterministic) sum of its points
template <typename T>
__global__ void k_segment_reduce(long long n, const unsigned *__restrict__ key, const unsigned *__restrict__ val,
                                 const double *__restrict__ y, long long plane, T *__restrict__ oty,
                                 T *__restrict__ cnt) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const unsigned k = key[j];
    if (k == 0xFFFFFFFFu) continue;
    if (j > 0 && key[j - 1] == k) continue;
    double s = 0.0;
    long long c = 0;
    for (long long t = j; t < n && key[t] == k; ++t) {
      s += y[val[t]];
      ++c;
    }
    oty[plane + k] = (T)s;
    cnt[plane + k] = (T)c;
  }
}

template <typename T>
__global__ void k_gather(long long n, const long long *__restrict__ vid, const T *__restrict__ theta_ghosted,
                         long long plane, long long z0, long long nz, double *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long lv = vid[i] - z0 * plane;
    out[i] = (lv >= 0 && lv < nz * plane) ? (double)theta_ghosted[plane + lv] : nan("");
  }
}

static inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

void launch_bin(int p, const DimTab &dt, long long n, const double *data, long long ld_point, long long ld_axis,
                const double *axes, long long *vid, unsigned *key, unsigned *val, cudaStream_t st) {
  if (n <= 0) return;
  k_bin<<<grid_for(n), 256, 0, st>>>(p, dt, n, data, ld_point, ld_axis, axes, vid, key, val);
  MVTV_CUDA(cudaGetLastError());
}

size_t sort_temp_bytes(long long n) {
  size_t bytes = 0;
  MVTV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned *)nullptr, (unsigned *)nullptr,
                                            (const unsigned *)nullptr, (unsigned *)nullptr, (int)n));
  return bytes;
}

void launch_sort(void *temp, size_t temp_bytes, long long n, const unsigned *key_in, unsigned *key_out,
                 const unsigned *val_in, unsigned *val_out, cudaStream_t st) {
  if (n <= 0) return;
  MVTV_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key_in, key_out, val_in, val_out, (int)n, 0, 32, st));
}

template <typename T>
void launch_segment_reduce(long long n, const unsigned *key, const unsigned *val, const double *y, long long plane,
                           T *oty, T *cnt, cudaStream_t st) {
  if (n <= 0) return;
  k_segment_reduce<T><<<grid_for(n), 256, 0, st>>>(n, key, val, y, plane, oty, cnt);
  MVTV_CUDA(cudaGetLastError());
}
template void launch_segment_reduce<double>(long long, const unsigned *, const unsigned *, const double *, long long,
                                            double *, double *, cudaStream_t);
template void launch_segment_reduce<float>(long long, const unsigned *, const unsigned *, const double *, long long,
                                           float *, float *, cudaStream_t);

template <typename T>
void launch_gather(long long n, const long long *vid, const T *theta_ghosted, long long plane, long long z0,
                   long long nz, double *out, cudaStream_t st) {
  if (n <= 0) return;
  k_gather<T><<<grid_for(n), 256, 0, st>>>(n, vid, theta_ghosted, plane, z0, nz, out);
  MVTV_CUDA(cudaGetLastError());
}
template void launch_gather<double>(long long, const long long *, const double *, long long, long long, long long,
                                    double *, cudaStream_t);
template void launch_gather<float>(long long, const long long *, const float *, long long, long long, long long,
                                   double *, cudaStream_t);

}  // namespace mvtv
