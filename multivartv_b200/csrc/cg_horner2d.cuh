// k_cg_horner2d: the Horner passes of the polynomial preconditioner of degree >= 2 on 2-D meshes (one GPU).
//
// z = P(A) D^-1 r with A = D^-1 M and P the degree-d Chebyshev polynomial (solver.cu, cheb_poly_coeffs).  Measured on 4096^2
// (profiles/r2b_probe2.log): 34.8 / 24.2 / 18.6 / 14.9 CG iterations per pass for d = 1..4 (Jacobi: 65.9) and 11.0 / 10.3 / 9.9 /
// 9.5 ms per pass.  P is evaluated in Horner form, one stencil pass per degree:
//     w_1 = c_d A z0 + c_{d-1} z0                      (FIRST: z0 = dinv r formed on the fly; reads r, dinv; writes w: 3 N;
//                                                       inside an x-update this pass is fused into k_cg_updprec2d)
//     w_k = A w_{k-1} + c_{d-k} z0 ,  k = 2..d         (reads w, dinv, r; writes w: 4 N;  z = w_d)
// with A w = w + rhoM dinv (K w - diag(K) w) (diag(c) is never read, as in k_cg_step2d).  The pass outputs rotate over z, the idle
// direction buffer and y (CgArgs::w_out_scr / w_in_scr), ending in z.  The last pass reduces r.z and commits it.
// Structure as k_cg_step2d: a warp owns a strip of 64 vertices, x-neighbours by shuffle, marching along the last axis.
#pragma once
#include "cg_step2d.cuh"

namespace mvtv {

template <typename T, int WARPS, int MINB, bool FIRST>
__global__ void __launch_bounds__(32 * WARPS, (MINB > 0 ? MINB : 1))
k_cg_horner2d(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
              const RedBuf rb, const int zchunk) {
  if (cg_done(a.S, a.rtol2)) return;
  const int it = (int)a.S[CS_ITERS];
  const T *__restrict__ w_in = cg_wsel(a, a.w_in_scr, it, false);
  T *__restrict__ w_out = cg_wsel(a, a.w_out_scr, it, false);
  const double c_lo = a.pc0, c_hi = a.pc1;
  const int final_pass = a.final_pass;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T *__restrict__ r = cg_rcur(a, it);
  const T *__restrict__ dinv = a.dinv;
  const T rhoM = (T)a.rhoM;
  const int m0 = (int)dt.m[0];                        // even, >= 2
  const int xw = blockIdx.x * (64 * WARPS) + warp * 64;
  const int x = xw + 2 * lane;
  const bool valid = x < m0;
  const int xo = valid ? x : m0 - 2;                  // out-of-mesh lanes replicate the last vertex (clamped neighbour)
  const bool edge = (lane == 0) || (lane == 31);
  const int xh = min((lane == 0) ? max(xw - 1, 0) : xw + 64, m0 - 1);
  const bool bx0 = (x == 0 || x == m0 - 1), bx1 = (x + 1 == 0 || x + 1 == m0 - 1);
  const int zc0 = blockIdx.y * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zfirst = zc0 - 1, zlast = zc1;            // one GPU: rows outside [0, nz) clamp to the boundary row

  T sa[2] = {T(0), T(0)}, sb[2] = {T(0), T(0)}, ha = T(0), hb = T(0);   // staged row: FIRST r, dinv ; else w (sb unused)
  T od[2] = {T(0), T(0)}, orr[2] = {T(0), T(0)};                        // !FIRST: dinv and r of an own row
  auto load_row = [&](int zz) {
    if (zz > zlast) return;
    const int zs = min(max(zz, 0), dt.nz - 1);
    const long long pb = (long long)(zs + 1) * dt.plane;
    if (FIRST) {
      ld2(r + pb + xo, sa);
      ld2(dinv + pb + xo, sb);
      if (edge) { ha = r[pb + xh]; hb = dinv[pb + xh]; }
    } else {
      ld2(w_in + pb + xo, sa);
      if (edge) ha = w_in[pb + xh];
      if (valid && zz >= zc0 && zz < zc1) {
        ld2(dinv + pb + x, od);
        ld2(r + pb + x, orr);
      }
    }
  };

  T A0[2] = {T(0), T(0)}, A1[2] = {T(0), T(0)}, A2[2] = {T(0), T(0)};
  T wcp[2] = {T(0), T(0)}, dcp[2] = {T(0), T(0)}, rcp[2] = {T(0), T(0)};   // w, dinv, r of the row retiring next
  double red[1] = {0.0};

  load_row(zfirst);
  for (int zz = zfirst; zz <= zlast; ++zz) {
    T v[2], hv, down[2], rown[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      v[k] = FIRST ? sb[k] * sa[k] : sa[k];
      down[k] = FIRST ? sb[k] : od[k];
      rown[k] = FIRST ? sa[k] : orr[k];
    }
    hv = FIRST ? hb * ha : ha;
    if (!valid) v[0] = v[1];
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
      const long long gz = dt.z0 + zz - 1;
      const bool bz = (gz == 0 || gz == dt.m[1] - 1);
      T outv[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool bxk = k ? bx1 : bx0;
        const T dk = (T)(bz ? (bxk ? st.diagK[3] : st.diagK[2]) : (bxk ? st.diagK[1] : st.diagK[0]));
        const T Aw = wcp[k] + rhoM * dcp[k] * (A0[k] - dk * wcp[k]);      // (D^-1 M w) at this vertex
        const T z0 = FIRST ? wcp[k] : dcp[k] * rcp[k];
        const T ov = FIRST ? (T)c_hi * Aw + (T)c_lo * z0 : Aw + (T)c_lo * z0;
        outv[k] = ov;
        red[0] += (double)rcp[k] * (double)ov;
      }
      st2(w_out + (long long)zz * dt.plane + x, outv[0], outv[1]);   // row zz-1 sits at (zz-1+1)*plane
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      A0[k] = A1[k];
      A1[k] = A2[k];
      A2[k] = T(0);
      wcp[k] = v[k];
      dcp[k] = down[k];
      rcp[k] = rown[k];
    }
  }
  if (!final_pass) return;
  double *S = a.S;
  grid_reduce<1, 1>(red, rb, [S](const double (&res)[1]) { cg_commit_rz(S, res); });
}

// coefficients c_0 .. c_d of P with 1 - t P(t) = T_{d+1}((theta - t)/delta) / T_{d+1}(theta/delta) on [bmax/kappa, bmax]
inline void cheb_poly_coeffs(int d, double bmax, double kappa, double *c /* d+1 */) {
  const double lo = bmax / kappa, th = 0.5 * (bmax + lo), de = 0.5 * (bmax - lo);
  // polynomials in t as coefficient arrays (degree <= d+1 <= 8)
  double Tm[10] = {1.0}, Tc[10] = {th / de, -1.0 / de}, Tn[10];
  for (int k = 1; k <= d; ++k) {   // T_{k+1} = 2 ((th - t)/de) T_k - T_{k-1}
    for (int i = 0; i < 10; ++i) Tn[i] = -Tm[i];
    for (int i = 0; i < 9; ++i) {
      Tn[i] += 2.0 * (th / de) * Tc[i];
      Tn[i + 1] += -2.0 / de * Tc[i];
    }
    for (int i = 0; i < 10; ++i) { Tm[i] = Tc[i]; Tc[i] = Tn[i]; }
  }
  const double t0 = Tc[0];         // T_{d+1}(theta/delta): value of the polynomial at t = 0
  for (int k = 0; k <= d; ++k) c[k] = -Tc[k + 1] / t0;   // P(t) = (1 - R(t))/t with R = T/t0, R(0) = 1
}

}  // namespace mvtv
