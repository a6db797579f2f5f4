// k_cg_updprec2d: CG vector update fused with the polynomial preconditioner, 2-D meshes, single GPU
// (the default on 2-D meshes with an even m0 on one GPU; 4096^2: 193 us against 137 + 87 us for the separate kernels,
// profiles/r2b_probe2.log).
//
// One CG iteration with MVTV_PRECOND_CHEB1 is   p = z + beta p, q = M p, p.q      (k_cg_step2d<STEP_Z>,   5 N words)
//                                               theta += a p, r -= a q, r.r       (k_cg_update,           6 N words)
//                                               z = P(D^-1 M) D^-1 r, r.z         (k_cg_step2d<STEP_PREC>, 3 N words)
// The last two share r: fused, r_new is formed in registers (also on the strip's halo element and on the chunk's two halo
// rows, which needs q there -- available on one GPU, where q is a full vector), fed straight into the preconditioner's
// stencil, and written once.  Reads theta, p, r, q, dinv; writes theta, r, z: 8 N words instead of 9 N, one launch and one
// grid reduction less per iteration.  r is updated OUT OF PLACE (CgArgs::r / r2 selected by the parity of the iterations): a CTA reads
// the rows next to its chunk as halo while their owner rewrites them.  theta is updated in place (no halo reads).
// Structure as k_cg_step2d: a warp owns a strip of 64 vertices (two per lane), x-neighbours by shuffle, march along the last
// axis with three accumulator sets in registers.  Scalars: alpha = r.z / p.q of the current parity; the epilogue commits
// {r.z, r.r} of the next parity and advances the iteration count (cg_commit_update); with a polynomial of degree >= 2 this is
// the first Horner pass (output buffer CgArgs::w_out_scr) and r.z comes from the last one.
#pragma once
#include "cg_step2d.cuh"

namespace mvtv {

template <int WARPS_, int MINB_ = 0>
struct Fused2dCfg {
  static constexpr int WARPS = WARPS_, MINB = MINB_, NT = 32 * WARPS_, TX = 64 * WARPS_;
};

template <typename T, typename Cfg>
__global__ void __launch_bounds__(Cfg::NT, (Cfg::MINB > 0 ? Cfg::MINB : 1))
k_cg_updprec2d(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
               const RedBuf rb, const int zchunk) {
  if (cg_done(a.S, a.rtol2)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int it = (int)a.S[CS_ITERS];
  const int cur = it & 1;
  const T alpha = (T)(a.S[2 * cur] / a.S[CS_PQ]);
  const T *__restrict__ r_in = cur ? a.r2 : a.r;     // cg_rcur
  T *__restrict__ r_out = cur ? a.r : a.r2;
  T *__restrict__ zo = cg_wsel(a, a.w_out_scr, it, true);   // first Horner pass of a degree >= 2 polynomial
  const T *__restrict__ p = a.pbuf[cur ^ 1];          // the direction k_cg_step2d<STEP_Z> just wrote
  const T *__restrict__ q = a.q;
  const T *__restrict__ dinv = a.dinv;
  T *__restrict__ x = a.x;
  const T rhoM = (T)a.rhoM;

  const int m0 = (int)dt.m[0];                        // even, >= 2
  const int xw = blockIdx.x * Cfg::TX + warp * 64;
  const int xx = xw + 2 * lane;
  const bool valid = xx < m0;
  const int xo = valid ? xx : m0 - 2;                 // out-of-mesh lanes replicate the last vertex (clamped neighbour)
  const bool edge = (lane == 0) || (lane == 31);
  const int xh = min((lane == 0) ? max(xw - 1, 0) : xw + 64, m0 - 1);
  const bool bx0 = (xx == 0 || xx == m0 - 1), bx1 = (xx + 1 == 0 || xx + 1 == m0 - 1);

  const int zc0 = blockIdx.y * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zfirst = zc0 - 1, zlast = zc1;            // single GPU: rows outside [0, nz) clamp to the boundary row

  // raw inputs of the next row, in flight while the current one is consumed
  T rr[2], rq[2], rd[2], rx[2], rp[2], hr = T(0), hq = T(0), hd = T(0);
  rr[0] = rr[1] = rq[0] = rq[1] = rd[0] = rd[1] = rx[0] = rx[1] = rp[0] = rp[1] = T(0);
  auto load_row = [&](int zz) {
    if (zz > zlast) return;
    const int zs = min(max(zz, 0), dt.nz - 1);
    const long long pb = (long long)(zs + 1) * dt.plane;
    ld2(r_in + pb + xo, rr);
    ld2(q + pb + xo, rq);
    ld2(dinv + pb + xo, rd);
    if (edge) {
      hr = r_in[pb + xh];
      hq = q[pb + xh];
      hd = dinv[pb + xh];
    }
    if (valid && zz >= zc0 && zz < zc1) {
      ld2(x + pb + xx, rx);
      ld2(p + pb + xx, rp);
    }
  };

  T A0[2] = {T(0), T(0)}, A1[2] = {T(0), T(0)}, A2[2] = {T(0), T(0)};
  T zcp[2] = {T(0), T(0)}, rcp[2] = {T(0), T(0)}, dcp[2] = {T(0), T(0)};   // z0, r_new, dinv of the row retiring next
  double red[2] = {0.0, 0.0};                                            // r.z, r.r

  load_row(zfirst);
  for (int zz = zfirst; zz <= zlast; ++zz) {
    const bool ownrow = zz >= zc0 && zz < zc1;
    // ---- r_new and z0 = dinv * r_new of row zz at the pair and at the strip's halo element
    T rn[2], v[2], down[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      rn[k] = rr[k] - alpha * rq[k];
      v[k] = rd[k] * rn[k];
      down[k] = rd[k];
    }
    const T hv = hd * (hr - alpha * hq);
    if (!valid) v[0] = v[1];
    if (ownrow && valid) {
      const long long ob = (long long)(zz + 1) * dt.plane + xx;
      st2(r_out + ob, rn[0], rn[1]);
      st2(x + ob, rx[0] + alpha * rp[0], rx[1] + alpha * rp[1]);
      red[1] += (double)rn[0] * (double)rn[0] + (double)rn[1] * (double)rn[1];
    }
    load_row(zz + 1);
    // ---- neighbours along axis 0 from the adjacent lanes; stencil contributions to output rows zz+1, zz, zz-1
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
    // ---- retire output row zz-1: z = pc0 z0 + pc1 dinv (M z0), with dinv*c = 1 - dinv*rhoM*diag(K)
    if (zz - 1 >= zc0 && valid) {
      const long long gz = dt.z0 + zz - 1;
      const bool bz = (gz == 0 || gz == dt.m[1] - 1);
      T outv[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool bxk = k ? bx1 : bx0;
        const T dk = (T)(bz ? (bxk ? st.diagK[3] : st.diagK[2]) : (bxk ? st.diagK[1] : st.diagK[0]));
        const T zv = (T)(a.pc0 + a.pc1) * zcp[k] + (T)a.pc1 * (rhoM * dcp[k] * (A0[k] - dk * zcp[k]));
        outv[k] = zv;
        red[0] += (double)rcp[k] * (double)zv;
      }
      st2(zo + (long long)zz * dt.plane + xx, outv[0], outv[1]);   // row zz-1 sits at (zz-1+1)*plane
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      A0[k] = A1[k];
      A1[k] = A2[k];
      A2[k] = T(0);
      zcp[k] = v[k];
      rcp[k] = rn[k];
      dcp[k] = down[k];
    }
  }
  double *S = a.S;
  const int fin = a.final_pass;
  grid_reduce<2, 2>(red, rb, [S, fin](const double (&res)[2]) {
    if (fin) cg_commit_update(S, res);
    else cg_commit_update_prec(S, res + 1);   // degree >= 2: r.z comes from the last Horner pass
  });
}

}  // namespace mvtv
