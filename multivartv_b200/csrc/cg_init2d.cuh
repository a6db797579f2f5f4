// k_cg_init2d: k_cg_init (kernels.cuh) for 2-D meshes in the marching / shuffle form of k_cg_step2d
// (the default on 2-D meshes with an even m0; 4096^2: 197 against 368 us, profiles/r2b_probe2.log).
//   b = Oty + rho*(D^T alpha + uscale*D^T u)   (never stored)      r = b - (diag(c) + rhoM D^T D) theta      theta_old = theta
//   r.z (z = dinv r), r.r, b.b                                      + the neighbours' ghost rows of r on several GPUs
// k_cg_init gathers the 9 stencil points of theta per vertex (0.45 of the HBM peak on 4096^2: the re-reads go through L1 / L2);
// here a warp owns a strip of 64 vertices, theta of a row is loaded once, its x-neighbours come by shuffle and the three rows of
// the stencil are accumulated in registers while marching along the last axis.  Algorithmic traffic as k_cg_init: 8 N words.
#pragma once
#include "cg_step2d.cuh"

namespace mvtv {

template <typename T, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
k_cg_init2d(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a, const RedBuf rb,
            const int zchunk) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T rhoM = (T)a.rhoM, rho = (T)a.rho, usc = (T)a.uscale;
  const int m0 = (int)dt.m[0];                        // even, >= 2
  const int xw = blockIdx.x * (64 * WARPS) + warp * 64;
  const int x = xw + 2 * lane;
  const bool valid = x < m0;
  const int xo = valid ? x : m0 - 2;                  // out-of-mesh lanes replicate the last vertex (clamped neighbour)
  const bool edge = (lane == 0) || (lane == 31);
  const int xh = min((lane == 0) ? max(xw - 1, 0) : xw + 64, m0 - 1);
  const int zc0 = blockIdx.y * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zlo = dt.has_lo ? -1 : 0;                 // theta's ghost rows hold the neighbours' rows (exchange_ghosts)
  const int zhi = dt.has_hi ? dt.nz : dt.nz - 1;
  const int zfirst = zc0 - 1, zlast = zc1;

  T tx[2] = {T(0), T(0)}, hx = T(0);                  // theta of the next row at the pair / at the strip's halo element
  T roty[2], rv1[2], rv2[2], rc[2], rd[2];            // own-row inputs of the row loaded last
  roty[0] = roty[1] = rv1[0] = rv1[1] = rv2[0] = rv2[1] = rc[0] = rc[1] = rd[0] = rd[1] = T(0);
  auto load_row = [&](int zz) {
    if (zz > zlast) return;
    const int zs = min(max(zz, zlo), zhi);
    const long long pb = (long long)(zs + 1) * dt.plane;
    ld2(a.x + pb + xo, tx);
    if (edge) hx = a.x[pb + xh];
    if (valid && zz >= zc0 && zz < zc1) {
      const long long ob = (long long)(zz + 1) * dt.plane + x;
      ld2(a.oty + ob, roty);
      ld2(a.v1 + ob, rv1);
      ld2(a.v2 + ob, rv2);
      ld2(a.c + ob, rc);
      ld2(a.dinv + ob, rd);
    }
  };

  T A0[2] = {T(0), T(0)}, A1[2] = {T(0), T(0)}, A2[2] = {T(0), T(0)};
  T xcp[2] = {T(0), T(0)}, bcp[2] = {T(0), T(0)}, ccp[2] = {T(0), T(0)}, dcp[2] = {T(0), T(0)};   // of the row retiring next
  double red[3] = {0.0, 0.0, 0.0};
  bool stored_peer = false;

  load_row(zfirst);
  for (int zz = zfirst; zz <= zlast; ++zz) {
    T v[2] = {tx[0], tx[1]};
    const T hv = hx;
    if (!valid) v[0] = v[1];
    T bown[2], cown[2], down[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      bown[k] = roty[k] + rho * (rv1[k] + usc * rv2[k]);
      cown[k] = rc[k];
      down[k] = rd[k];
    }
    load_row(zz + 1);
    T left = __shfl_up_sync(0xffffffffu, v[1], 1);
    T right = __shfl_down_sync(0xffffffffu, v[0], 1);
    if (lane == 0) left = hv;
    if (lane == 31) right = hv;
    const T W[2][3] = {{left, v[0], v[1]}, {v[0], v[1], right}};
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        A2[k] += (T)st.coef[dx] * W[k][dx];
        A1[k] += (T)st.coef[dx + 3] * W[k][dx];
        A0[k] += (T)st.coef[dx + 6] * W[k][dx];
      }
    if (zz - 1 >= zc0 && valid) {   // retire row zz-1
      const long long ob = (long long)zz * dt.plane + x;
      T rvv[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const T rv = bcp[k] - (ccp[k] * xcp[k] + rhoM * A0[k]);
        const T zv = rv * dcp[k];
        rvv[k] = rv;
        red[0] += (double)rv * (double)zv;
        red[1] += (double)rv * (double)rv;
        red[2] += (double)bcp[k] * (double)bcp[k];
      }
      st2(a.r + ob, rvv[0], rvv[1]);
      st2(a.xold + ob, xcp[0], xcp[1]);
      if (a.peer) {  // fill the neighbours' ghost rows of r
        if (zz - 1 == 0 && a.peer->has_lo) { st2((T *)a.peer->rghost_at_prev + x, rvv[0], rvv[1]); stored_peer = true; }
        if (zz - 1 == dt.nz - 1 && a.peer->has_hi) { st2((T *)a.peer->rghost_at_next + x, rvv[0], rvv[1]); stored_peer = true; }
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      A0[k] = A1[k];
      A1[k] = A2[k];
      A2[k] = T(0);
      xcp[k] = v[k];
      bcp[k] = bown[k];
      ccp[k] = cown[k];
      dcp[k] = down[k];
    }
  }
  if (stored_peer) __threadfence_system();
  double *S = a.S, *raw = a.raw;
  const PeerTab *peer = a.peer;
  const unsigned long long sr = a.seq_red, sh = a.seq_halo;
  const int fold = a.fold;
  grid_reduce<3, 3>(red, rb, [S, raw, peer, sr, sh, fold](const double (&res)[3]) {
    if (peer) {
      __threadfence_system();
      if (peer->has_lo) st_release_sys(peer->hflag_at_prev, sh);
      if (peer->has_hi) st_release_sys(peer->hflag_at_next, sh);
      peer_post(*peer, sr, res, 3);
      if (fold) {   // what k_cg_peer_commit_init does
        double v[3];
        peer_wait_sum(*peer, sr, v, 3);
        cg_commit_init(S, v);
      }
    } else if (raw) { raw[0] = res[0]; raw[1] = res[1]; raw[2] = res[2]; }
    else cg_commit_init(S, res);
  });
}

}  // namespace mvtv
