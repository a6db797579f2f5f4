// Internal declarations shared by the CUDA translation units of libmvtv_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "../../include/mvtv.h"

#define MVTV_MAXK 15    // 2^MAXP - 1 difference blocks
#define MVTV_MAXSUB 16  // subsets of a block's axis set

namespace mvtv {

// ---------------------------------------------------------------------------------------------
// Mesh / operator table, passed BY VALUE as a __grid_constant__ kernel parameter (< 4 KB), so
// several plans with different meshes can coexist in one process.
//
// Vertex vectors live in a "ghosted slab" layout: plane = prod(m[0..P-2]) values per plane of the
// last axis, nz owned planes [z0, z0+nz) of the global last axis, plus one ghost plane below and one
// above: element (q, zl) of the slab is at  (zl+1)*plane + q.  On one GPU z0=0, nz=m[P-1] and the
// ghost planes are never read (boundary guards).  The dual variable u is stored PADDED: block b,
// row owned by vertex v is at  b*usz + idx(v)  (rows that do not exist at the high boundary of an
// axis in the block's axis set are simply never touched) -- so a row and its vertex share an address
// pattern and every access is as coalesced as the vertex vectors are.
// ---------------------------------------------------------------------------------------------
struct DimTab {
  int P;
  int nz;       // owned planes of the last axis
  int has_lo;   // 1 if a neighbour rank owns plane z0-1 (the ghost plane below holds real data)
  int has_hi;   // 1 if a neighbour rank owns plane z0+nz
  long long m[MVTV_MAXP];       // GLOBAL mesh dims
  long long stride[MVTV_MAXP];  // stride[a] = prod(m[0..a-1]); stride[P-1] == plane
  long long plane, z0, Nloc;    // Nloc = plane*nz
  long long usz;                // plane*(nz+2): size of one ghosted vertex vector / one u block
};

// Difference blocks of D in create_D order (cpp-code/utils.cpp:245-269).
struct BlockTab {
  int K;
  int mask[MVTV_MAXK];          // effective axis set S' of block b (bit a = difference along axis a)
  int nsub[MVTV_MAXK];          // 2^|S'|
  double scale[MVTV_MAXK];      // c_S
  int sub[MVTV_MAXK][MVTV_MAXSUB];        // the subsets e of S' (sub[b][0] == 0)
  long long off[MVTV_MAXK][MVTV_MAXSUB];  // sum_{a in e} stride[a]
};

// Reference (compact) row layout of D, used only at the ABI boundary (pack / unpack of u, apply_D).
struct RowTab {
  int K;
  int mask[MVTV_MAXK];
  long long rows[MVTV_MAXK];     // rows of block b on the GLOBAL mesh
  long long row_off[MVTV_MAXK];  // offset of block b in the stacked row order
  long long rstride[MVTV_MAXK][MVTV_MAXP];  // strides of the reduced (m_a - [a in S']) tensor
  long long R;
};

// 3^P-point stencil of K = D^T D with Neumann (clamped-index) boundaries:
//   (K x)[i] = sum_o coef[o] * x[clamp(i + o)],  o in {-1,0,1}^P, digit a of o = o_a+1
struct StencilTab {
  int npts;  // 3^P
  double coef[81];
  double diagK[16];  // diag(K) per boundary class: bit a set = vertex is at a boundary of axis a
};

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &msg);

#define MVTV_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      throw ::mvtv::Error(MVTV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)

#define MVTV_REQUIRE(cond, msg)                                      \
  do {                                                               \
    if (!(cond)) throw ::mvtv::Error(MVTV_ERR_INVALID, (msg));       \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Deterministic grid-wide reduction: every block stores NR partials, the last block to finish adds
// them in a fixed order (so results do not depend on block scheduling) and ONE thread of it runs the
// epilogue functor epi(res) with the NR totals.  That thread runs after every block of the grid has
// finished its loads, so the epilogue may overwrite scalars the kernel read at entry.
// Slots [0,NSUM) are sums, slots [NSUM,NR) are maxima.
// ---------------------------------------------------------------------------------------------
struct RedBuf {
  double *partials;        // nblocks*NR
  unsigned int *counter;   // self-resetting ticket
};

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// blockDim.x*blockDim.y*blockDim.z must be a multiple of 32 and <= 1024.
template <int NR, int NSUM, typename Epi>
__device__ __forceinline__ void grid_reduce(double (&v)[NR], const RedBuf &rb, Epi epi) {
  __shared__ double s_part[32][NR];
  __shared__ bool s_last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    double w = (k < NSUM) ? warp_sum(v[k]) : warp_max(v[k]);
    if (lane == 0) s_part[warp][k] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      double w = (lane < nwarp) ? s_part[lane][k] : ((k < NSUM) ? 0.0 : -1.0e300);
      w = (k < NSUM) ? warp_sum(w) : warp_max(w);
      if (lane == 0) rb.partials[(size_t)bid * NR + k] = w;
    }
    if (lane == 0) {
      __threadfence();
      unsigned t = atomicInc(rb.counter, nblocks - 1);
      s_last = (t == nblocks - 1);
    }
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double acc[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[k] = (k < NSUM) ? 0.0 : -1.0e300;
    for (unsigned b = tid; b < nblocks; b += nthr) {
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        double w = __ldcg(&rb.partials[(size_t)b * NR + k]);
        acc[k] = (k < NSUM) ? acc[k] + w : fmax(acc[k], w);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      double w = (k < NSUM) ? warp_sum(acc[k]) : warp_max(acc[k]);
      if (lane == 0) s_part[warp][k] = w;
    }
    __syncthreads();
    if (warp == 0) {
      double res[NR];
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        double w = (lane < nwarp) ? s_part[lane][k] : ((k < NSUM) ? 0.0 : -1.0e300);
        res[k] = (k < NSUM) ? warp_sum(w) : warp_max(w);
      }
      if (lane == 0) epi(res);
    }
  }
}
#endif  // __CUDACC__

}  // namespace mvtv
