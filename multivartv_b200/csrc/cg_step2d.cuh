// k_cg_step2d: the fused CG direction update + SpMV of k_cg_step (kernels.cuh) specialised for 2-D meshes,
// WITHOUT shared memory.  Same three modes, same arguments, same reduction epilogue, same ghost-row protocol.
//
// A warp owns a strip of 64 consecutive vertices of axis 0 (two per lane, one 16-byte load / store per array and
// row) and marches along axis 1 over a chunk of rows.  The 3x3 clamped stencil of diag(c) + rhoM*D^T D needs, per
// row, the lane's two values of p_new and one neighbour on each side: those come from the adjacent lanes by warp
// shuffle; lane 0 / lane 31 load the strip's halo element themselves (a second, 8-byte load).  Reuse along the
// marching axis stays in registers exactly as in k_cg_step (row z adds to the accumulators of output rows z-1, z,
// z+1).  HBM latency is covered by a register ring of PF rows in flight per lane instead of a cp.async ring, so
// there is no barrier in the marching loop and no LDGSTS / STS / LDS round trip per element.
// Requires an even m0 (16-byte alignment of every row); the host falls back to k_cg_step otherwise.
// STEP_PREC can also run without reading diag(c) at all (Step2dCfg::NOC): with dinv = 1/(c + rhoM*diag(K)),
// dinv*(M z0) = z0 + rhoM*dinv*(K z0 - diag(K) z0), so the kernel reads r and dinv and writes z: 3 N words.
//
// Algorithmic traffic: as k_cg_step (JACOBI 6 N, Z 5 N, PREC 4 N words).
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace mvtv {

template <typename T> struct Vec2T;
template <> struct Vec2T<double> { using type = double2; };
template <> struct Vec2T<float> { using type = float2; };

template <typename T>
__device__ __forceinline__ void ld2(const T *__restrict__ p, T (&v)[2]) {
  const typename Vec2T<T>::type w = *reinterpret_cast<const typename Vec2T<T>::type *>(p);
  v[0] = w.x;
  v[1] = w.y;
}
template <typename T>
__device__ __forceinline__ void st2(T *__restrict__ p, T a, T b) {
  typename Vec2T<T>::type w;
  w.x = a;
  w.y = b;
  *reinterpret_cast<typename Vec2T<T>::type *>(p) = w;
}

// WARPS per CTA, PF rows in flight per lane, VPL vertices per lane (2 or 4: one or two 16-byte accesses per array and
// row), MINB = minimum resident CTAs per SM asked of the compiler (0: none), NOC: STEP_PREC derives diag(c) from dinv
// instead of reading it (dinv*c = 1 - dinv*rhoM*diag(K)), i.e. 3 N words instead of 4 N.
// DKSEL: NOC picks diag(K) of a vertex with selects on constant-bank operands instead of holding 2*VPL values in registers.
// IDX32: element offsets inside the ghosted slab in 32 bits (a slab holds < 2^31 vertices, enforced at plan creation; the
// two ghost planes keep the largest offset below 2^32), which saves the registers of the 64-bit index arithmetic.
template <int WARPS_, int PF_, int VPL_ = 2, int MINB_ = 0, bool NOC_ = false, bool DKSEL_ = false, bool IDX32_ = false>
struct Step2dCfg {
  static constexpr int WARPS = WARPS_, PF = PF_, VPL = VPL_, NG = VPL_ / 2, MINB = MINB_;
  static constexpr bool NOC = NOC_, DKSEL = DKSEL_, IDX32 = IDX32_;
  static constexpr int NT = 32 * WARPS_, SW = 32 * VPL_, TX = SW * WARPS_;
  static_assert(VPL_ == 2 || VPL_ == 4, "two or four vertices per lane");
};

template <typename T, typename Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::NT, (Cfg::MINB > 0 ? Cfg::MINB : 1))
k_cg_step2d(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
            const RedBuf rb, const int zchunk) {
  if (cg_done(a.S, a.rtol2)) return;
  constexpr int PF = Cfg::PF, VPL = Cfg::VPL, NG = Cfg::NG, SW = Cfg::SW;
  constexpr bool NOC = Cfg::NOC && (MODE == STEP_PREC);
  using idx_t = typename std::conditional<Cfg::IDX32, unsigned, long long>::type;
  const idx_t plane = (idx_t)dt.plane;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int it = (int)a.S[CS_ITERS];
  const int cur = it & 1;
  const bool first = (MODE == STEP_PREC) ? true : (it == 0);   // "first": no p_old term
  const T beta = first ? T(0) : (T)(a.S[2 * cur] / a.S[2 * (cur ^ 1)]);
  const T *__restrict__ p_in = a.pbuf[cur];
  T *__restrict__ p_out = a.pbuf[cur ^ 1];
  const T *__restrict__ rr = (MODE == STEP_Z) ? a.z : cg_rcur(a, it);
  T *__restrict__ zo = cg_wsel(a, a.w_out_scr, it, false);   // STEP_PREC output (first Horner pass of a degree >= 2 polynomial)
  const T *__restrict__ dinv = a.dinv;
  const T rhoM = (T)a.rhoM;

  const int m0 = (int)dt.m[0];                       // even, >= 2
  const int xw = blockIdx.x * Cfg::TX + warp * SW;   // first vertex of this warp's strip
  const int x = xw + VPL * lane;                     // this lane's vertices x .. x+VPL-1, in NG aligned pairs
  bool valid[NG];
  int xo[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    valid[g] = x + 2 * g < m0;
    xo[g] = valid[g] ? x + 2 * g : m0 - 2;           // out-of-mesh pairs replicate the last vertex (clamped neighbour)
  }
  const bool edge = (lane == 0) || (lane == 31);
  const int xh = min((lane == 0) ? max(xw - 1, 0) : xw + SW, m0 - 1);   // the strip's halo element of this lane

  const int zc0 = blockIdx.y * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zlo = dt.has_lo ? -1 : 0;
  const int zhi = dt.has_hi ? dt.nz : dt.nz - 1;
  const int zfirst = zc0 - 1, zlast = zc1;
  if (a.peer) {  // the neighbours fill our ghost rows of r (z) directly: wait for the version this launch needs
    if (tid == 0) {
      const unsigned long long need = (MODE == STEP_Z) ? a.seq_zhalo : a.seq_halo;
      const unsigned long long *fp = (MODE == STEP_Z) ? a.peer->zflag_from_prev : a.peer->hflag_from_prev;
      const unsigned long long *fn = (MODE == STEP_Z) ? a.peer->zflag_from_next : a.peer->hflag_from_next;
      if (zc0 == 0 && dt.has_lo) peer_spin(fp, need, a.peer->error);
      if (zc1 == dt.nz && dt.has_hi) peer_spin(fn, need, a.peer->error);
    }
    __syncthreads();
  }
  // NOC: diag(K) of this lane's vertices for an interior row and for a boundary row of the marching axis
  constexpr bool DKSEL = Cfg::DKSEL;
  T dKi[VPL], dKb[VPL];
  bool bxk[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const bool bx = (x + k == 0 || x + k == m0 - 1);
    bxk[k] = bx;
    dKi[k] = (NOC && !DKSEL) ? (T)(bx ? st.diagK[1] : st.diagK[0]) : T(0);
    dKb[k] = (NOC && !DKSEL) ? (T)(bx ? st.diagK[3] : st.diagK[2]) : T(0);
  }

  // register ring: raw inputs of PF rows in flight
  T ra[PF][VPL], rb_[PF][VPL], rc[PF][VPL], rcc[PF][VPL];   // r|z, dinv, p_old at the lane's vertices; diag(c) of owned rows
  T ha[PF], hb[PF], hc[PF];                                 // the same at the strip's halo element (lanes 0 and 31)
#pragma unroll
  for (int s = 0; s < PF; ++s) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) ra[s][k] = rb_[s][k] = rc[s][k] = rcc[s][k] = T(0);
    ha[s] = hb[s] = hc[s] = T(0);
  }
  auto load_row = [&](int zz, int s) {
    if (zz > zlast) return;
    const int zs = min(max(zz, zlo), zhi);
    const idx_t pb = (idx_t)(zs + 1) * plane;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      ld2(rr + (pb + (idx_t)xo[g]), *reinterpret_cast<T(*)[2]>(&ra[s][2 * g]));
      if (MODE != STEP_Z) ld2(dinv + (pb + (idx_t)xo[g]), *reinterpret_cast<T(*)[2]>(&rb_[s][2 * g]));
      if (!first) ld2(p_in + (pb + (idx_t)xo[g]), *reinterpret_cast<T(*)[2]>(&rc[s][2 * g]));
    }
    if (edge) {
      ha[s] = rr[pb + (idx_t)xh];
      if (MODE != STEP_Z) hb[s] = dinv[pb + (idx_t)xh];
      if (!first) hc[s] = p_in[pb + (idx_t)xh];
    }
    if (!NOC && zz >= zc0 && zz < zc1) {
#pragma unroll
      for (int g = 0; g < NG; ++g)
        if (valid[g]) ld2(a.c + ((idx_t)(zz + 1) * plane + (idx_t)(x + 2 * g)), *reinterpret_cast<T(*)[2]>(&rcc[s][2 * g]));
    }
  };

  T A0[VPL], A1[VPL], A2[VPL], pcp[VPL], cqp[VPL], rcp[VPL], dcp[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) A0[k] = A1[k] = A2[k] = pcp[k] = cqp[k] = rcp[k] = dcp[k] = T(0);
  double red[1] = {0.0};

#pragma unroll
  for (int s = 0; s < PF; ++s) load_row(zfirst + s, s);
  for (int zb = zfirst; zb <= zlast; zb += PF) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      const int zz = zb + s;
      if (zz <= zlast) {
        // ---- combine: p_new of row zz at the lane's vertices and at the halo element
        T v[VPL], hv, rown[VPL], down[VPL], cown[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          v[k] = (MODE == STEP_Z) ? ra[s][k] : rb_[s][k] * ra[s][k];
          if (!first) v[k] += beta * rc[s][k];
          rown[k] = ra[s][k];
          down[k] = rb_[s][k];
          cown[k] = rcc[s][k];
        }
        hv = (MODE == STEP_Z) ? ha[s] : hb[s] * ha[s];
        if (!first) hv += beta * hc[s];
#pragma unroll
        for (int g = 0; g < NG; ++g)
          if (!valid[g]) v[2 * g] = v[2 * g + 1];            // replicate vertex m0-1
        load_row(zz + PF, s);                                // the slot is free: next row of the ring goes in flight
        {
          const int zs = min(max(zz, zlo), zhi);
          const bool own = (zz == zs) && ((zz >= zc0 && zz < zc1) || (zz < 0 && zc0 == 0) || (zz >= dt.nz && zc1 == dt.nz));
          if (MODE != STEP_PREC && own) {
#pragma unroll
            for (int g = 0; g < NG; ++g)
              if (valid[g]) st2(p_out + ((idx_t)(zs + 1) * plane + (idx_t)(x + 2 * g)), v[2 * g], v[2 * g + 1]);
          }
        }
        // ---- neighbours along axis 0 from the adjacent lanes
        T left = __shfl_up_sync(0xffffffffu, v[VPL - 1], 1);
        T right = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (lane == 0) left = hv;
        if (lane == 31) right = hv;
        // ---- stencil contributions of row zz to output rows zz+1 (A2), zz (A1), zz-1 (A0)
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const T W[3] = {k == 0 ? left : v[k > 0 ? k - 1 : 0], v[k], k == VPL - 1 ? right : v[k < VPL - 1 ? k + 1 : 0]};
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            A2[k] += (T)st.coef[dx] * W[dx];
            A1[k] += (T)st.coef[dx + 3] * W[dx];
            A0[k] += (T)st.coef[dx + 6] * W[dx];
          }
        }
        // ---- retire output row zz-1
        if (zz - 1 >= zc0) {
          const idx_t ob = (idx_t)zz * plane + (idx_t)x;       // local row zz-1 sits at (zz-1+1)*plane
          const long long gz = dt.z0 + zz - 1;
          const bool bz = (gz == 0 || gz == dt.m[1] - 1);
          T outv[VPL];
#pragma unroll
          for (int k = 0; k < VPL; ++k) {
            const T pv = pcp[k];
            if (MODE == STEP_PREC) {
              T zv;
              if (NOC) {   // dinv*q = z0 + dinv*rhoM*(K z0 - diag(K) z0): diag(c) never read
                const T dk = DKSEL ? (T)(bz ? (bxk[k] ? st.diagK[3] : st.diagK[2]) : (bxk[k] ? st.diagK[1] : st.diagK[0]))
                                   : (bz ? dKb[k] : dKi[k]);
                zv = (T)(a.pc0 + a.pc1) * pv + (T)a.pc1 * (rhoM * dcp[k] * (A0[k] - dk * pv));
              } else {
                const T qv = cqp[k] * pv + rhoM * A0[k];
                zv = (T)a.pc0 * pv + (T)a.pc1 * (dcp[k] * qv);
              }
              outv[k] = zv;
              if (valid[k / 2]) red[0] += (double)rcp[k] * (double)zv;
            } else {
              const T qv = cqp[k] * pv + rhoM * A0[k];
              outv[k] = qv;
              if (valid[k / 2]) red[0] += (double)pv * (double)qv;
            }
          }
#pragma unroll
          for (int g = 0; g < NG; ++g)
            if (valid[g]) {
              if (MODE == STEP_PREC) {
                st2(zo + ob + 2 * g, outv[2 * g], outv[2 * g + 1]);
                if (a.peer) {  // fill the neighbours' ghost rows of z
                  if (zz - 1 == 0 && dt.has_lo) { st2((T *)a.peer->zghost_at_prev + x + 2 * g, outv[2 * g], outv[2 * g + 1]); __threadfence_system(); }
                  if (zz - 1 == dt.nz - 1 && dt.has_hi) { st2((T *)a.peer->zghost_at_next + x + 2 * g, outv[2 * g], outv[2 * g + 1]); __threadfence_system(); }
                }
              } else {
                st2(a.q + ob + 2 * g, outv[2 * g], outv[2 * g + 1]);
              }
            }
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          A0[k] = A1[k];
          A1[k] = A2[k];
          A2[k] = T(0);
          pcp[k] = v[k];
          cqp[k] = cown[k];
          rcp[k] = rown[k];
          dcp[k] = down[k];
        }
      }
    }
  }
  double *S = a.S, *raw = a.raw;
  const PeerTab *peer = a.peer;
  const unsigned long long sr = a.seq_red, sz = a.seq_zhalo;
  const int fold = a.fold;
  if (MODE == STEP_PREC && !a.final_pass) return;    // first pass of a degree >= 2 polynomial (one GPU): r.z comes from the last pass
  grid_reduce<1, 1>(red, rb, [S, raw, peer, sr, sz, fold](const double (&res)[1]) {
    if (peer) {
      if (MODE == STEP_PREC) {
        __threadfence_system();
        if (peer->has_lo) st_release_sys(peer->zflag_at_prev, sz);
        if (peer->has_hi) st_release_sys(peer->zflag_at_next, sz);
      }
      peer_post(*peer, sr, res, 1);
      if (fold) {   // what k_cg_peer_commit_rz / k_cg_peer_commit_pq do
        double v[1];
        peer_wait_sum(*peer, sr, v, 1);
        if (MODE == STEP_PREC) cg_commit_rz(S, v);
        else S[CS_PQ] = v[0];
      }
    } else if (raw) raw[0] = res[0];
    else if (MODE == STEP_PREC) cg_commit_rz(S, res);
    else S[CS_PQ] = res[0];
  });
}

}  // namespace mvtv
