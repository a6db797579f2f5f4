// Host-side launchers of setup.cu (kept in a separate translation unit: CUB is slow to compile).
#pragma once
#include "mvtv_internal.cuh"

namespace mvtv {
// data element of point i, axis a: data[i*ld_point + a*ld_axis]  (column-major: 1, n ; row-major: p, 1)
void launch_bin(int p, const DimTab &dt, long long n, const double *data, long long ld_point, long long ld_axis,
                const double *axes, long long *vid, unsigned *key, unsigned *val, cudaStream_t st);
size_t sort_temp_bytes(long long n);
void launch_sort(void *temp, size_t temp_bytes, long long n, const unsigned *key_in, unsigned *key_out,
                 const unsigned *val_in, unsigned *val_out, cudaStream_t st);
template <typename T>
void launch_segment_reduce(long long n, const unsigned *key, const unsigned *val, const double *y, long long plane,
                           T *oty, T *cnt, cudaStream_t st);
template <typename T>
void launch_gather(long long n, const long long *vid, const T *theta_ghosted, long long plane, long long z0,
                   long long nz, double *out, cudaStream_t st);
}  // namespace mvtv
