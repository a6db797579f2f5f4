// k_cg_step3dh: "hybrid" variant of the fused CG direction update + SpMV for 3-D meshes
// (EXPERIMENTAL: opt-in with MVTV_STEP3D=hyb; logic-checked on the CPU SIMT emulator, not yet run on a GPU).
//
// k_cg_step stages the raw tiles of r, dinv, p_old through a cp.async ring in shared memory, combines them there and reads
// a (RY+2) x 3 window per thread: two barriers and ~70 instructions per vertex and plane.  k_cg_step3d keeps everything in
// registers, which costs 175-255 registers and re-reads the y-halo rows.  Here the two are crossed:
//   * a warp owns ONE row of the plane tile (64 vertices of axis 0, two per lane); a CTA is TY own rows + 2 halo rows;
//   * the raw inputs of a row go global -> registers with 16-byte loads, PF planes ahead (no ring in shared memory);
//   * p_new of the row is formed in registers, written back (own rows) and stored ONCE into a small shared plane
//     [2 buffers][TY+2 rows][64 + 2 halo elements]; ONE barrier per plane (the plane is double buffered);
//   * every own-row warp reads its 3 x 4 window (rows y-1, y, y+1; x-1 .. x+2) from the shared plane and adds the plane's
//     contribution to the accumulators of output planes z-1, z, z+1, as k_cg_step does.
// Shared memory: 2*(TY+2)*68*sizeof(T) (8.7 KB for TY = 6 in fp64), so occupancy is set by registers (~100).
// STEP_PREC derives diag(c) from dinv like k_cg_step2d (3 N words).  Requires an even m0.
#pragma once
#include "cg_step2d.cuh"

namespace mvtv {

template <int TY_, int PF_ = 1, bool NOC_ = true>
struct Step3dhCfg {
  static constexpr int TY = TY_, PF = PF_, NW = TY_ + 2, NT = 32 * (TY_ + 2), TX = 64, ROW = 68;
  static constexpr bool NOC = NOC_;
};

template <typename T, typename Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::NT)
k_cg_step3dh(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
             const RedBuf rb, const int zchunk) {
  if (cg_done(a.S, a.rtol2)) return;
  constexpr int TY = Cfg::TY, PF = Cfg::PF, NW = Cfg::NW, ROW = Cfg::ROW;
  constexpr bool NOC = Cfg::NOC && (MODE == STEP_PREC);
  __shared__ __align__(16) T sp[2][NW][ROW];   // p_new of the plane: [1] left halo, [2..65] the strip, [66] right halo
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
  const int xw = bx * Cfg::TX;
  const int x = xw + 2 * lane;
  const bool xvalid = x < m0;
  const int xo = xvalid ? x : m0 - 2;                 // out-of-mesh lanes replicate the last vertex (clamped neighbour)
  const bool edge = (lane == 0) || (lane == 31);
  const int xh = min((lane == 0) ? max(xw - 1, 0) : xw + 64, m0 - 1);
  const int yrow = by * TY - 1 + warp;                // warp 0 and warp TY+1 hold the clamped halo rows
  const int yc = min(max(yrow, 0), m1 - 1);
  const bool ownrow = warp >= 1 && warp <= TY && yrow < m1;
  const bool valid = ownrow && xvalid;
  const int rowoff = yc * m0 + xo, rowoff_h = yc * m0 + xh;
  const long long ownoff = (long long)yc * m0 + x;

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
  const int clsxy0 = ((x == 0 || x == m0 - 1) ? 1 : 0) | ((yc == 0 || yc == m1 - 1) ? 2 : 0);
  const int clsxy1 = ((x + 1 == 0 || x + 1 == m0 - 1) ? 1 : 0) | ((yc == 0 || yc == m1 - 1) ? 2 : 0);

  // register ring: raw inputs of PF planes in flight
  T ra[PF][2], rb_[PF][2], rc[PF][2], rcc[PF][2], ha[PF], hb[PF], hc[PF];
#pragma unroll
  for (int s = 0; s < PF; ++s) {
    ra[s][0] = ra[s][1] = rb_[s][0] = rb_[s][1] = rc[s][0] = rc[s][1] = rcc[s][0] = rcc[s][1] = T(0);
    ha[s] = hb[s] = hc[s] = T(0);
  }
  auto load_plane = [&](int zz, int s) {
    if (zz > zlast) return;
    const int zs = min(max(zz, zlo), zhi);
    const long long pb = (long long)(zs + 1) * dt.plane;
    ld2(rr + pb + rowoff, ra[s]);
    if (MODE != STEP_Z) ld2(dinv + pb + rowoff, rb_[s]);
    if (!first) ld2(p_in + pb + rowoff, rc[s]);
    if (edge) {
      ha[s] = rr[pb + rowoff_h];
      if (MODE != STEP_Z) hb[s] = dinv[pb + rowoff_h];
      if (!first) hc[s] = p_in[pb + rowoff_h];
    }
    if (!NOC && valid && zz >= zc0 && zz < zc1) ld2(a.c + (long long)(zz + 1) * dt.plane + ownoff, rcc[s]);
  };

  T A0[2] = {T(0), T(0)}, A1[2] = {T(0), T(0)}, A2[2] = {T(0), T(0)};
  T pcp[2] = {T(0), T(0)}, cqp[2] = {T(0), T(0)}, rcp[2] = {T(0), T(0)}, dcp[2] = {T(0), T(0)};
  double red[1] = {0.0};

#pragma unroll
  for (int s = 0; s < PF; ++s) load_plane(zfirst + s, s);
  for (int zb = zfirst; zb <= zlast; zb += PF) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      const int zz = zb + s;
      if (zz <= zlast) {
        // ---- combine: p_new of this warp's row in plane zz (pair + the strip's halo element)
        T v[2], hv;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          v[k] = (MODE == STEP_Z) ? ra[s][k] : rb_[s][k] * ra[s][k];
          if (!first) v[k] += beta * rc[s][k];
        }
        hv = (MODE == STEP_Z) ? ha[s] : hb[s] * ha[s];
        if (!first) hv += beta * hc[s];
        if (!xvalid) v[0] = v[1];                          // replicate vertex m0-1
        const T rown[2] = {ra[s][0], ra[s][1]}, down[2] = {rb_[s][0], rb_[s][1]}, cown[2] = {rcc[s][0], rcc[s][1]};
        load_plane(zz + PF, s);                            // the slot is free: the next plane of the ring goes in flight
        T(*plane)[ROW] = sp[(zz - zfirst) & 1];
        st2(&plane[warp][2 + 2 * lane], v[0], v[1]);
        if (lane == 0) plane[warp][1] = hv;
        if (lane == 31) plane[warp][66] = hv;
        {
          const int zs = min(max(zz, zlo), zhi);
          const bool own = (zz == zs) && ((zz >= zc0 && zz < zc1) || (zz < 0 && zc0 == 0) || (zz >= dt.nz && zc1 == dt.nz));
          if (MODE != STEP_PREC && own && valid) st2(p_out + (long long)(zs + 1) * dt.plane + ownoff, v[0], v[1]);
        }
        __syncthreads();   // the plane is complete; the other buffer is free again once everybody passed the previous barrier
        if (ownrow) {
          // ---- stencil contributions of plane zz to output planes zz+1 (A2), zz (A1), zz-1 (A0): rows warp-1, warp, warp+1
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const T *row = plane[warp - 1 + dy];
            T mid[2];
            ld2(row + 2 + 2 * lane, mid);
            const T W[2][3] = {{row[1 + 2 * lane], mid[0], mid[1]}, {mid[0], mid[1], row[4 + 2 * lane]}};
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int ci = dx + 3 * dy;
                A2[k] += (T)st.coef[ci] * W[k][dx];
                A1[k] += (T)st.coef[ci + 9] * W[k][dx];
                A0[k] += (T)st.coef[ci + 18] * W[k][dx];
              }
          }
          // ---- retire output plane zz-1
          if (zz - 1 >= zc0 && valid) {
            const long long ob = (long long)zz * dt.plane + ownoff;   // plane zz-1 sits at (zz-1+1)*plane
            const long long gz = dt.z0 + zz - 1;
            const int bz = (gz == 0 || gz == dt.m[2] - 1) ? 4 : 0;
            T outv[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const T pv = pcp[k];
              if (MODE == STEP_PREC) {
                T zv;
                if (NOC) {   // dinv*q = z0 + dinv*rhoM*(K z0 - diag(K) z0): diag(c) never read
                  const T dk = (T)st.diagK[(k ? clsxy1 : clsxy0) | bz];
                  zv = (T)(a.pc0 + a.pc1) * pv + (T)a.pc1 * (rhoM * dcp[k] * (A0[k] - dk * pv));
                } else {
                  const T qv = cqp[k] * pv + rhoM * A0[k];
                  zv = (T)a.pc0 * pv + (T)a.pc1 * (dcp[k] * qv);
                }
                outv[k] = zv;
                red[0] += (double)rcp[k] * (double)zv;
              } else {
                const T qv = cqp[k] * pv + rhoM * A0[k];
                outv[k] = qv;
                red[0] += (double)pv * (double)qv;
              }
            }
            if (MODE == STEP_PREC) {
              st2(a.z + ob, outv[0], outv[1]);
              if (a.peer) {  // fill the neighbours' ghost planes of z
                if (zz - 1 == 0 && dt.has_lo) { st2((T *)a.peer->zghost_at_prev + ownoff, outv[0], outv[1]); __threadfence_system(); }
                if (zz - 1 == dt.nz - 1 && dt.has_hi) { st2((T *)a.peer->zghost_at_next + ownoff, outv[0], outv[1]); __threadfence_system(); }
              }
            } else {
              st2(a.q + ob, outv[0], outv[1]);
            }
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
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
