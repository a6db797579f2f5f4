// k_cg_step3d: the shared-memory-free variant of k_cg_step for 3-D meshes (EXPERIMENTAL: opt-in with MVTV_STEP3D=shfl,
// written after round 1's GPU minutes were spent -- compiled for sm_100a but NOT yet run on a GPU; the default 3-D path
// stays k_cg_step).  Same three modes, arguments, reduction epilogue and ghost-plane protocol as k_cg_step /
// k_cg_step2d (cg_step2d.cuh, whose 2-D form is measured: 0.92 of the HBM peak for STEP_Z).
//
// A warp owns 64 consecutive vertices of axis 0 (two per lane, 16-byte accesses) times RY consecutive rows of axis 1 and
// marches along axis 2.  Per plane a lane loads its pair on RY + 2 rows (the two extra rows are the clamped y-neighbours:
// they are the own rows of the adjacent warp of the same CTA, so they come out of L1 / L2), forms p_new on all of them,
// gets the x-neighbours from the adjacent lanes by shuffle (lanes 0 / 31 load the strip's halo element), and adds the
// plane's contribution to the three accumulator sets of output planes z-1, z, z+1 (27-point clamped stencil).
// STEP_PREC derives diag(c) from dinv like k_cg_step2d (3 N words).
#pragma once
#include "cg_step2d.cuh"

namespace mvtv {

template <int WARPS_, int RY_, int MINB_ = 0, bool NOC_ = true>
struct Step3dCfg {
  static constexpr int WARPS = WARPS_, RY = RY_, MINB = MINB_, NT = 32 * WARPS_, TX = 64, TY = RY_ * WARPS_;
  static constexpr bool NOC = NOC_;
};

template <typename T, typename Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::NT, (Cfg::MINB > 0 ? Cfg::MINB : 1))
k_cg_step3d(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
            const RedBuf rb, const int zchunk) {
  if (cg_done(a.S, a.rtol2)) return;
  constexpr int RY = Cfg::RY, NR = Cfg::RY + 2;
  constexpr bool NOC = Cfg::NOC && (MODE == STEP_PREC);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int it = (int)a.S[CS_ITERS];
  const int cur = it & 1;
  const bool first = (MODE == STEP_PREC) ? true : (it == 0);   // "first": no p_old term
  const T beta = first ? T(0) : (T)(a.S[2 * cur] / a.S[2 * (cur ^ 1)]);
  const T *__restrict__ p_in = a.pbuf[cur];
  T *__restrict__ p_out = a.pbuf[cur ^ 1];
  const T *__restrict__ rr = (MODE == STEP_Z) ? a.z : a.r;
  const T *__restrict__ dinv = a.dinv;
  const T rhoM = (T)a.rhoM;

  const int m0 = (int)dt.m[0], m1 = (int)dt.m[1];     // m0 even, >= 2
  const int ntx = (m0 + Cfg::TX - 1) / Cfg::TX;
  const int bx = blockIdx.x % ntx, by = blockIdx.x / ntx;
  const int xw = bx * Cfg::TX;                        // first vertex of the strip (all warps of a CTA share it)
  const int x = xw + 2 * lane;
  const bool xvalid = x < m0;
  const int xo = xvalid ? x : m0 - 2;                 // out-of-mesh lanes replicate the last vertex (clamped neighbour)
  const bool edge = (lane == 0) || (lane == 31);
  const int xh = min((lane == 0) ? max(xw - 1, 0) : xw + 64, m0 - 1);
  const int y0 = by * Cfg::TY + warp * RY;            // first own row of this warp
  // loaded rows r = 0 .. RY+1 are y0-1 .. y0+RY, clamped (Neumann); own rows are r = 1 .. RY
  int rowoff[NR], rowoff_h[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int yy = min(max(y0 - 1 + r, 0), m1 - 1);
    rowoff[r] = yy * m0 + xo;
    rowoff_h[r] = yy * m0 + xh;
  }
  bool valid[RY];
#pragma unroll
  for (int j = 0; j < RY; ++j) valid[j] = xvalid && (y0 + j < m1);

  const int zc0 = blockIdx.y * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zlo = dt.has_lo ? -1 : 0;
  const int zhi = dt.has_hi ? dt.nz : dt.nz - 1;
  const int zfirst = zc0 - 1, zlast = zc1;
  if (a.peer) {  // the neighbours fill our ghost planes of r (z) directly: wait for the version this launch needs
    if (tid == 0) {
      const unsigned long long need = (MODE == STEP_Z) ? a.seq_zhalo : a.seq_halo;
      const unsigned long long *fp = (MODE == STEP_Z) ? a.peer->zflag_from_prev : a.peer->hflag_from_prev;
      const unsigned long long *fn = (MODE == STEP_Z) ? a.peer->zflag_from_next : a.peer->hflag_from_next;
      if (zc0 == 0 && dt.has_lo) peer_spin(fp, need, a.peer->error);
      if (zc1 == dt.nz && dt.has_hi) peer_spin(fn, need, a.peer->error);
    }
    __syncthreads();
  }
  // NOC: boundary class bits of this lane's outputs within a plane (bit 0: axis 0, bit 1: axis 1)
  int cls[RY][2];
#pragma unroll
  for (int j = 0; j < RY; ++j)
#pragma unroll
    for (int k = 0; k < 2; ++k)
      cls[j][k] = ((x + k == 0 || x + k == m0 - 1) ? 1 : 0) | ((y0 + j == 0 || y0 + j == m1 - 1) ? 2 : 0);

  // raw inputs of the next plane, in flight while the current one is consumed
  T ra[NR][2], rb_[NR][2], rc[NR][2], rcc[RY][2], ha[NR], hb[NR], hc[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    ra[r][0] = ra[r][1] = rb_[r][0] = rb_[r][1] = rc[r][0] = rc[r][1] = T(0);
    ha[r] = hb[r] = hc[r] = T(0);
  }
#pragma unroll
  for (int j = 0; j < RY; ++j) rcc[j][0] = rcc[j][1] = T(0);
  auto load_plane = [&](int zz) {
    if (zz > zlast) return;
    const int zs = min(max(zz, zlo), zhi);
    const long long pb = (long long)(zs + 1) * dt.plane;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      ld2(rr + pb + rowoff[r], ra[r]);
      if (MODE != STEP_Z) ld2(dinv + pb + rowoff[r], rb_[r]);
      if (!first) ld2(p_in + pb + rowoff[r], rc[r]);
      if (edge) {
        ha[r] = rr[pb + rowoff_h[r]];
        if (MODE != STEP_Z) hb[r] = dinv[pb + rowoff_h[r]];
        if (!first) hc[r] = p_in[pb + rowoff_h[r]];
      }
    }
    if (!NOC && zz >= zc0 && zz < zc1) {
#pragma unroll
      for (int j = 0; j < RY; ++j)
        if (valid[j]) ld2(a.c + (long long)(zz + 1) * dt.plane + (long long)(y0 + j) * m0 + x, rcc[j]);
    }
  };

  T A0[RY][2], A1[RY][2], A2[RY][2], pcp[RY][2], cqp[RY][2], rcp[RY][2], dcp[RY][2];
#pragma unroll
  for (int j = 0; j < RY; ++j)
#pragma unroll
    for (int k = 0; k < 2; ++k) A0[j][k] = A1[j][k] = A2[j][k] = pcp[j][k] = cqp[j][k] = rcp[j][k] = dcp[j][k] = T(0);
  double red[1] = {0.0};

  load_plane(zfirst);
  for (int zz = zfirst; zz <= zlast; ++zz) {
    // ---- combine: p_new of plane zz on the RY+2 rows, at the pair and at the strip's halo element
    T v[NR][2], hv[NR], rown[RY][2], down[RY][2], cown[RY][2];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        v[r][k] = (MODE == STEP_Z) ? ra[r][k] : rb_[r][k] * ra[r][k];
        if (!first) v[r][k] += beta * rc[r][k];
      }
      hv[r] = (MODE == STEP_Z) ? ha[r] : hb[r] * ha[r];
      if (!first) hv[r] += beta * hc[r];
      if (!xvalid) v[r][0] = v[r][1];                    // replicate vertex m0-1
    }
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        rown[j][k] = ra[j + 1][k];
        down[j][k] = rb_[j + 1][k];
        cown[j][k] = rcc[j][k];
      }
    load_plane(zz + 1);                                  // the registers are free: the next plane goes in flight
    {
      const int zs = min(max(zz, zlo), zhi);
      const bool own = (zz == zs) && ((zz >= zc0 && zz < zc1) || (zz < 0 && zc0 == 0) || (zz >= dt.nz && zc1 == dt.nz));
      if (MODE != STEP_PREC && own) {
#pragma unroll
        for (int j = 0; j < RY; ++j)
          if (valid[j]) st2(p_out + (long long)(zs + 1) * dt.plane + (long long)(y0 + j) * m0 + x, v[j + 1][0], v[j + 1][1]);
      }
    }
    // ---- stencil contributions of plane zz to output planes zz+1 (A2), zz (A1), zz-1 (A0)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      T left = __shfl_up_sync(0xffffffffu, v[r][1], 1);
      T right = __shfl_down_sync(0xffffffffu, v[r][0], 1);
      if (lane == 0) left = hv[r];
      if (lane == 31) right = hv[r];
      const T W[2][3] = {{left, v[r][0], v[r][1]}, {v[r][0], v[r][1], right}};
      // loaded row r is the dy = r - j neighbour of own row j (dy in 0..2)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int j = r - dy;
        if (j >= 0 && j < RY) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const int ci = dx + 3 * dy;
              A2[j][k] += (T)st.coef[ci] * W[k][dx];
              A1[j][k] += (T)st.coef[ci + 9] * W[k][dx];
              A0[j][k] += (T)st.coef[ci + 18] * W[k][dx];
            }
        }
      }
    }
    // ---- retire output plane zz-1
    if (zz - 1 >= zc0) {
      const long long gz = dt.z0 + zz - 1;
      const int bz = (gz == 0 || gz == dt.m[2] - 1) ? 4 : 0;
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        if (!valid[j]) continue;
        const long long ob = (long long)zz * dt.plane + (long long)(y0 + j) * m0 + x;   // plane zz-1 sits at (zz-1+1)*plane
        T outv[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const T pv = pcp[j][k];
          if (MODE == STEP_PREC) {
            T zv;
            if (NOC) {   // dinv*q = z0 + dinv*rhoM*(K z0 - diag(K) z0): diag(c) never read
              const T dk = (T)st.diagK[cls[j][k] | bz];
              zv = (T)(a.pc0 + a.pc1) * pv + (T)a.pc1 * (rhoM * dcp[j][k] * (A0[j][k] - dk * pv));
            } else {
              const T qv = cqp[j][k] * pv + rhoM * A0[j][k];
              zv = (T)a.pc0 * pv + (T)a.pc1 * (dcp[j][k] * qv);
            }
            outv[k] = zv;
            red[0] += (double)rcp[j][k] * (double)zv;
          } else {
            const T qv = cqp[j][k] * pv + rhoM * A0[j][k];
            outv[k] = qv;
            red[0] += (double)pv * (double)qv;
          }
        }
        if (MODE == STEP_PREC) {
          st2(a.z + ob, outv[0], outv[1]);
          if (a.peer) {  // fill the neighbours' ghost planes of z
            const long long q = (long long)(y0 + j) * m0 + x;
            if (zz - 1 == 0 && dt.has_lo) { st2((T *)a.peer->zghost_at_prev + q, outv[0], outv[1]); __threadfence_system(); }
            if (zz - 1 == dt.nz - 1 && dt.has_hi) { st2((T *)a.peer->zghost_at_next + q, outv[0], outv[1]); __threadfence_system(); }
          }
        } else {
          st2(a.q + ob, outv[0], outv[1]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        A0[j][k] = A1[j][k];
        A1[j][k] = A2[j][k];
        A2[j][k] = T(0);
        pcp[j][k] = v[j + 1][k];
        cqp[j][k] = cown[j][k];
        rcp[j][k] = rown[j][k];
        dcp[j][k] = down[j][k];
      }
  }
  double *S = a.S, *raw = a.raw;
  const PeerTab *peer = a.peer;
  const unsigned long long sr = a.seq_red, sz = a.seq_zhalo;
  const int fold = a.fold;
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
