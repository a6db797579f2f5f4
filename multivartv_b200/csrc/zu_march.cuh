// k_zu_march: the fused z/u update in scatter form, marching along the last mesh axis.
//
// Replaces, in one pass over HBM (cpp-code/solvers.cpp:115,117-120,113 ; rcpp solvers.cpp:112,114-122):
//   alpha = softthresh(D*theta - u, lambda/rho)      primal_residual = alpha - D*theta      u += primal_residual
//   D^T alpha, D^T u_new, D^T u_old  (-> b of the next x-update and the dual residual)
//   ||r||^2, ||s||^2, ||D^T u||^2, ||D theta||^2, ||alpha||^2, max|theta - theta_old|
//
// Work decomposition.  A CTA owns an in-plane tile of vertices and marches over a chunk of planes of the
// last axis.  The thread of vertex v evaluates ONLY the rows v owns (one per difference block): D*theta from
// two shared-memory planes of theta, u_old from registers prefetched one plane ahead, writes u_new, and
// scatters sign*scale*{alpha, u_new, u_old} to the 2^P vertices v+f the row touches.  The scatter is
// folded per offset class f: contributions with f_z = 1 stay in registers and are consumed when the march
// reaches the next plane ("carry"); contributions with an in-plane offset go once through shared memory.
// Every row is therefore evaluated exactly once per chunk (the gather form in kernels.cuh re-derives each
// row 2^|S| times), at the price of one low-side halo column/row of threads per tile.
// Block axis sets are compile-time (BlockMasks<P, V>), so all per-offset arrays live in registers.
//
// Algorithmic traffic: read u_old (R), theta, theta_old ; write u_new (R), D^T alpha, D^T u  = T*(2R + 4N).
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace mvtv {

enum { ZV_REFERENCE = 0, ZV_INTENDED = 1, ZV_P1 = 2 };  // block tables known at compile time

// axis set S' of block b in create_D order (cpp-code/utils.cpp:245-269, :187 for the quirk); must agree with
// mvtv_plan::build_tables (checked on the host before the kernel is selected)
__host__ __device__ constexpr int zu_num_blocks(int P, int V) { return V == ZV_P1 ? 1 : (1 << P) - 1; }
__host__ __device__ constexpr int zu_block_mask(int P, int V, int b) {
  if (V == ZV_P1) return 1;
  const int K = (1 << P) - 1;
  const int num = (b == 0) ? K : b;
  int S = 0;
  for (int a = 0; a < P; ++a)
    if ((num >> (P - 1 - a)) & 1) S |= 1 << a;
  int cnt = 0, lowest = -1;
  for (int a = 0; a < P; ++a)
    if ((S >> a) & 1) {
      if (lowest < 0) lowest = a;
      ++cnt;
    }
  if (cnt > 1 && V == ZV_REFERENCE) S = (S & ~(1 << lowest)) | 1;
  return S;
}

__host__ __device__ constexpr int zu_popc(int x) {
  int c = 0;
  for (int a = 0; a < 8; ++a) c += (x >> a) & 1;
  return c;
}

template <int N, typename F, int I = 0>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<N, F, I + 1>(static_cast<F &&>(f));
  }
}

// EX x EY x EW = extended (owned + one low-side halo) in-plane tile = threads of the CTA
template <int P_, int EX_, int EY_, int EW_>
struct ZuCfg {
  static constexpr int P = P_, Q = P_ - 1;
  static constexpr int EX = EX_, EY = (Q >= 2 ? EY_ : 1), EW = (Q >= 3 ? EW_ : 1);
  static constexpr int NT = EX * EY * EW;
  static constexpr int OX = EX - 1, OY = (Q >= 2 ? EY - 1 : 1), OW = (Q >= 3 ? EW - 1 : 1);  // owned extents
  static constexpr int SX = EX + 1, SY = (Q >= 2 ? EY + 1 : 1), SW = (Q >= 3 ? EW + 1 : 1);  // theta tile (+1 high)
  static constexpr int ST = SX * SY * SW;
  static constexpr int NS = (ST + NT - 1) / NT;
  static constexpr int NF = (1 << Q);           // in-plane offset classes
  static constexpr int SMEM_ELEMS = 3 * ST + (NF - 1) * 3 * NT;
};

template <typename T, typename Cfg, int V>
__global__ void __launch_bounds__(Cfg::NT)
k_zu_march(const __grid_constant__ DimTab dt, const __grid_constant__ BlockTab bt, const ZuArgs<T> a,
           const RedBuf rb, const int zchunk) {
  constexpr int P = Cfg::P, Q = Cfg::Q, EX = Cfg::EX, EY = Cfg::EY, NT = Cfg::NT;
  constexpr int OX = Cfg::OX, OY = Cfg::OY, OW = Cfg::OW, SX = Cfg::SX, SY = Cfg::SY, ST = Cfg::ST, NS = Cfg::NS;
  constexpr int NF = Cfg::NF, K = zu_num_blocks(P, V);
  constexpr int ZBIT = 1 << Q;
  MVTV_DYN_SMEM(smem_raw);
  T *sth = reinterpret_cast<T *>(smem_raw);  // [3][ST] theta planes zz, zz+1 and (in flight) zz+2; slot = plane % 3
  T *sh = sth + 3 * ST;                      // [(NF-1)][3][NT] in-plane exchange

  const int tid = threadIdx.x;
  const int m0 = (int)dt.m[0];
  const int m1 = (Q >= 2) ? (int)dt.m[1] : 1;
  const int m2 = (Q >= 3) ? (int)dt.m[2] : 1;
  const long long mz = dt.m[Q];
  const int ntx = (m0 + OX - 1) / OX;
  const int nty = (m1 + OY - 1) / OY;
  int bid = blockIdx.x;
  const int bx = bid % ntx;
  bid /= ntx;
  const int by = bid % nty;
  const int bw = bid / nty;
  // extended tile origin = owned origin - 1 on every in-plane axis
  const int x0 = bx * OX - 1, y0 = (Q >= 2) ? by * OY - 1 : 0, w0 = (Q >= 3) ? bw * OW - 1 : 0;

  // this thread's vertex
  const int ex = tid % EX, ey = (tid / EX) % EY, ew = tid / (EX * EY);
  const int gx = x0 + ex, gy = y0 + ey, gw = w0 + ew;
  const bool inmesh = gx >= 0 && gx < m0 && gy >= 0 && gy < m1 && gw >= 0 && gw < m2;
  bool owned_xy = inmesh && ex >= 1;
  if (Q >= 2) owned_xy = owned_xy && ey >= 1;
  if (Q >= 3) owned_xy = owned_xy && ew >= 1;
  const long long qoff = inmesh ? gx + (long long)m0 * (gy + (long long)m1 * gw) : 0;
  // in-plane validity of rows: hi_ok bit a <=> i_a + 1 < m_a
  int hi_ok = 0;
  hi_ok |= (gx + 1 < m0) << 0;
  if (Q >= 2) hi_ok |= (gy + 1 < m1) << 1;
  if (Q >= 3) hi_ok |= (gw + 1 < m2) << 2;
  // position of this vertex in the theta tile (tile origin = extended origin)
  const int sidx = ex + SX * (ey + SY * ew);

  // theta tile loader: clamped sources (values outside the mesh are never used by a valid row)
  int tsrc[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const int e = tid + k * NT;
    const int sx = e % SX, sy = (e / SX) % SY, sw = e / (SX * SY);
    const int cx = min(max(x0 + sx, 0), m0 - 1);
    const int cy = min(max(y0 + sy, 0), m1 - 1);
    const int cw = min(max(w0 + sw, 0), m2 - 1);
    tsrc[k] = cx + m0 * (cy + m1 * cw);
  }

  const int zc0 = blockIdx.y * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zlo = dt.has_lo ? -1 : 0;
  const int zhi = dt.has_hi ? dt.nz : dt.nz - 1;
  const T kappa = (T)a.kappa, usc = (T)a.uscale;

  auto load_theta = [&](int z) {  // cp.async: plane clamp(z) -> slot (z+3) % 3; always commits a group
    const int zs = min(max(z, zlo), zhi);
    const long long pb = (long long)(zs + 1) * dt.plane;
    T *dst = sth + ((z + 3) % 3) * ST;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const int e = tid + k * NT;
      if (e < ST) cp_async<sizeof(T)>(dst + e, a.theta + pb + tsrc[k]);
    }
    cp_async_commit();
  };
  // u_old of this thread's rows in plane z (prefetched one plane ahead)
  auto load_u = [&](int z, T (&dst)[K]) {
    const bool zvalid = inmesh && z >= zlo && z <= zhi && (dt.z0 + z) >= 0 && (dt.z0 + z) < mz;
    const long long pb = (long long)(z + 1) * dt.plane + qoff;
    static_for<K>([&](auto bc) {
      constexpr int b = decltype(bc)::value;
      dst[b] = zvalid ? a.u_old[(size_t)b * dt.usz + pb] : T(0);
    });
  };

  double red[ZR_N] = {0, 0, 0, 0, 0, 0};
  T carry[NF][3];  // contributions of the previous plane's rows to this plane's vertices (f_z = 1), per in-plane class
#pragma unroll
  for (int f = 0; f < NF; ++f) carry[f][0] = carry[f][1] = carry[f][2] = T(0);

  // first plane processed: zc0-1 (pre-step, only to build the carry) unless it lies below the global mesh
  const int zstart = (dt.z0 + zc0 - 1 >= 0) ? zc0 - 1 : zc0;
  T ucur[K], unext[K];
  T thp_cur = T(0), thp_next = T(0);                  // theta_old of this vertex, prefetched like u
  load_theta(zstart);
  load_theta(zstart + 1);
  load_u(zstart, ucur);
  if (a.theta_prev && owned_xy && zstart >= zc0) thp_cur = a.theta_prev[(long long)(zstart + 1) * dt.plane + qoff];
  for (int zz = zstart; zz < zc1; ++zz) {
    const bool pre = zz < zc0;                        // pre-step: rows are owned by the previous chunk / rank
    const bool ghostrow = pre && zz < 0;              // ... except the ghost plane's rows, kept up to date here
    load_theta(zz + 2);
    if (zz + 1 < zc1) {
      load_u(zz + 1, unext);
      if (a.theta_prev && owned_xy) thp_next = a.theta_prev[(long long)(zz + 2) * dt.plane + qoff];
    }
    cp_async_wait<1>();  // this thread's copies of planes <= zz+1 have landed
    __syncthreads();     // ... everybody's; previous plane's exchange buffer fully consumed
    const T *t0 = sth + ((zz + 3) % 3) * ST;
    const T *t1 = sth + ((zz + 4) % 3) * ST;
    const long long gz = dt.z0 + zz;
    const bool z_hi_ok = gz + 1 < mz;
    const long long pb = (long long)(zz + 1) * dt.plane + qoff;
    const T th_c = t0[sidx];   // read now: after the exchange barrier the slot may already be refilled

    T H[NF][3], CN[NF][3];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      H[f][0] = carry[f][0]; H[f][1] = carry[f][1]; H[f][2] = carry[f][2];
      CN[f][0] = CN[f][1] = CN[f][2] = T(0);
    }
    static_for<K>([&](auto bc) {
      constexpr int b = decltype(bc)::value;
      constexpr int S = zu_block_mask(P, V, b);
      constexpr int Sxy = S & (ZBIT - 1);
      constexpr bool hasz = (S & ZBIT) != 0;
      if (pre && !hasz) return;                       // only f_z = 1 contributions leave a pre-step plane
      const bool valid = inmesh && ((Sxy & ~hi_ok) == 0) && (!hasz || z_hi_ok);
      if (valid) {
        const T sc = (T)bt.scale[b];
        // D*theta at the row: sum over subsets f of S of sign(f) theta[v + f]
        T d = T(0);
        static_for<(1 << P)>([&](auto fc) {
          constexpr int f = decltype(fc)::value;
          if constexpr ((f & ~S) == 0) {
            constexpr int fxy = f & (ZBIT - 1);
            constexpr int off = (fxy & 1) + ((fxy >> 1) & 1) * SX + ((fxy >> 2) & 1) * SX * SY;
            const T t = ((f & ZBIT) ? t1 : t0)[sidx + off];
            if constexpr ((zu_popc(f) & 1) != 0) d -= t; else d += t;
          }
        });
        d *= sc;
        const T uo = usc * ucur[b];
        T al, un, pr;
        if (a.init) {
          al = d; un = uo; pr = T(0);
        } else {
          al = soft_threshold<T>(d - uo, kappa);      // cpp :117
          pr = al - d;                                // cpp :119
          un = uo + pr;                               // cpp :120
        }
        const bool mine = (!pre && owned_xy) || (ghostrow && owned_xy);
        if (mine && !a.init) a.u_new[(size_t)b * dt.usz + pb] = un;
        if (!pre && owned_xy) {
          red[ZR_R2] += (double)pr * (double)pr;
          red[ZR_DTH2] += (double)d * (double)d;
          red[ZR_AL2] += (double)al * (double)al;
        }
        const T w1 = sc * al, w2 = sc * un, w3 = sc * uo;
        static_for<(1 << P)>([&](auto fc) {
          constexpr int f = decltype(fc)::value;
          if constexpr ((f & ~S) == 0) {
            constexpr int fxy = f & (ZBIT - 1);
            constexpr bool neg = (zu_popc(f) & 1) != 0;
            if constexpr ((f & ZBIT) != 0) {
              CN[fxy][0] += neg ? -w1 : w1; CN[fxy][1] += neg ? -w2 : w2; CN[fxy][2] += neg ? -w3 : w3;
            } else {
              H[fxy][0] += neg ? -w1 : w1; H[fxy][1] += neg ? -w2 : w2; H[fxy][2] += neg ? -w3 : w3;
            }
          }
        });
      }
    });
    if (!pre) {
      // in-plane exchange: vertex v receives H[f] of the thread at v - f
#pragma unroll
      for (int f = 1; f < NF; ++f)
#pragma unroll
        for (int q = 0; q < 3; ++q) sh[((f - 1) * 3 + q) * NT + tid] = H[f][q];
      __syncthreads();
      if (owned_xy) {
        T t1s = H[0][0], t2s = H[0][1], t3s = H[0][2];
#pragma unroll
        for (int f = 1; f < NF; ++f) {
          const int nb = tid - (f & 1) - ((f >> 1) & 1) * EX - ((f >> 2) & 1) * EX * EY;
          t1s += sh[((f - 1) * 3 + 0) * NT + nb];
          t2s += sh[((f - 1) * 3 + 1) * NT + nb];
          t3s += sh[((f - 1) * 3 + 2) * NT + nb];
        }
        a.v1[pb] = t1s;
        a.v2[pb] = t2s;
        const double sv = (a.mode == MVTV_MODE_RCPP) ? (double)t2s - (double)t3s : (double)t1s + (double)t3s;
        red[ZR_S2] += sv * sv;
        red[ZR_DTU2] += (double)t2s * (double)t2s;
        if (a.theta_prev) red[ZR_DMAX] = fmax(red[ZR_DMAX], fabs((double)th_c - (double)thp_cur));
      }
    } else {
      __syncthreads();  // nobody may refill a theta slot while the pre-step's rows are still being evaluated
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      carry[f][0] = CN[f][0]; carry[f][1] = CN[f][1]; carry[f][2] = CN[f][2];
    }
    static_for<K>([&](auto bc) {
      constexpr int b = decltype(bc)::value;
      ucur[b] = unext[b];
    });
    thp_cur = thp_next;
  }
  cp_async_wait<0>();
  double *out = a.red_out;
  grid_reduce<ZR_N, ZR_NSUM>(red, rb, [out](const double (&res)[ZR_N]) {
#pragma unroll
    for (int k = 0; k < ZR_N; ++k) out[k] = res[k];
  });
}

}  // namespace mvtv
