// k_cg_step3d: the shared-memory-free CG kernels for 3-D meshes with an even m0 (the default there; measured on 512^3:
// STEP_Z 1.02 ms = 0.81 of the HBM peak against 1.12 ms for the shared-memory ring of k_cg_step, STEP_PREC 0.78 against
// 1.02 ms, profiles/r2_call1_step3d_probe.log).  Same modes, arguments, reduction epilogue and ghost-plane protocol as
// k_cg_step / k_cg_step2d, plus the two modes of the polynomial preconditioner of degree >= 2 and of the fused update:
//   STEP_HORNER   w_out = D^-1 M w_in + pc0 * D^-1 r          (pass k >= 2 of the Horner form; reads w_in, dinv, r: 4 N words)
//   STEP_INIT     the CG initialisation of k_cg_init in marching form: r = b - M theta, theta_old = theta, r.z (z = D^-1 r), r.r,
//                 b.b with b = Oty + rho (D^T alpha + uscale D^T u) never stored (8 N words; theta's halo rows come out of L1 / L2
//                 instead of 27 gathers per vertex)
//   STEP_UPDPREC  theta += alpha p ; r_new = r - alpha q (out of place) ; w_out = pc0 z0 + pc1 D^-1 M z0, z0 = D^-1 r_new
//                 (k_cg_update fused with the first preconditioner pass: reads theta, p, r, q, dinv, writes theta, r, w:
//                 8 N words instead of 6 N + 3 N; r_new is also formed on the halo rows / planes, which needs q there)
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

// TMAD_: 0 = the rows of a plane are loaded straight into registers, one plane ahead of their use; >= 2 = TMA staging: lanes
// issue one bulk copy (cp.async.bulk, 512 bytes) per row and array into a per-warp ring of TMAD_ stages in shared memory,
// TMAD_ - 1 planes ahead, tracked by one mbarrier per stage.  The ring is private to the warp (its own lanes are the only
// readers, so a stage is free again as soon as the warp has passed it: no empty-barrier, no CTA barrier in the loop); the
// halo rows it duplicates with the neighbouring warps come out of L2.
template <int WARPS_, int RY_, int MINB_ = 0, bool NOC_ = true, int TMAD_ = 0>
struct Step3dCfg {
  static constexpr int WARPS = WARPS_, RY = RY_, MINB = MINB_, NT = 32 * WARPS_, TX = 64, TY = RY_ * WARPS_;
  static constexpr bool NOC = NOC_;
  static constexpr int TMAD = TMAD_;
  // arrays staged with their halo rows per mode: JACOBI r, dinv, p ; Z z, p ; PREC r, dinv ; HORNER w ; UPDPREC r, dinv, q ; INIT theta
  __host__ __device__ static constexpr int narr(int mode) { return (mode == 0 || mode == 4) ? 3 : ((mode == 1 || mode == 2) ? 2 : 1); }
  __host__ __device__ static constexpr size_t smem_bytes(int mode, size_t esz) {
    return TMAD_ ? (size_t)WARPS_ * TMAD_ * (narr(mode) * (RY_ + 2) * 64 * esz + 8) : 0;
  }
};

template <typename T, typename Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::NT, (Cfg::MINB > 0 ? Cfg::MINB : 1))
k_cg_step3d(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
            const RedBuf rb, const int zchunk) {
  if (MODE != STEP_INIT && cg_done(a.S, a.rtol2)) return;   // STEP_INIT starts a solve: the scalars are the previous solve's
  constexpr int RY = Cfg::RY, NR = Cfg::RY + 2;
  constexpr bool POLY = (MODE == STEP_PREC || MODE == STEP_HORNER || MODE == STEP_UPDPREC);   // writes a preconditioner pass
  constexpr bool NOC = (Cfg::NOC && MODE == STEP_PREC) || MODE == STEP_HORNER || MODE == STEP_UPDPREC;
  constexpr bool STAGE_B = (MODE == STEP_JACOBI || MODE == STEP_PREC || MODE == STEP_UPDPREC);   // dinv on all loaded rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int it = (int)a.S[CS_ITERS];
  const int cur = it & 1;
  const bool first = (POLY || MODE == STEP_INIT) ? true : (it == 0);   // "first": no p_old term
  const bool stage_c = (MODE == STEP_UPDPREC) || !first;   // third staged array: p_old (JACOBI, Z) or q (UPDPREC)
  const T beta = first ? T(0) : (T)(a.S[2 * cur] / a.S[2 * (cur ^ 1)]);
  const T alpha = (MODE == STEP_UPDPREC) ? (T)(a.S[2 * cur] / a.S[CS_PQ]) : T(0);
  const T *__restrict__ p_in = (MODE == STEP_UPDPREC) ? a.q : a.pbuf[cur];
  T *__restrict__ p_out = a.pbuf[cur ^ 1];
  const T *__restrict__ rcur = cg_rcur(a, it);
  const T *__restrict__ w_in = cg_wsel(a, a.w_in_scr, it, false);
  const T *__restrict__ rr = (MODE == STEP_Z) ? a.z : ((MODE == STEP_HORNER) ? w_in : ((MODE == STEP_INIT) ? a.x : rcur));
  T *__restrict__ r_out = (MODE == STEP_UPDPREC) ? (cur ? a.r : a.r2) : nullptr;
  const T *__restrict__ p_dir = a.pbuf[cur ^ 1];   // STEP_UPDPREC: the direction the step kernel just wrote
  T *__restrict__ zo = cg_wsel(a, a.w_out_scr, it, MODE == STEP_UPDPREC);
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

  // chunk of the marching axis this CTA owns.  CTAs are scheduled in blockIdx order: on several GPUs the two chunks that
  // wait for a neighbour's ghost planes go last, so that no CTA spins on a flag while interior work is still queued
  const int nchunk = (int)gridDim.y;
  int chunk = (int)blockIdx.y;
  if (a.peer && nchunk > 2) chunk = (chunk < nchunk - 2) ? chunk + 1 : (chunk == nchunk - 2 ? 0 : nchunk - 1);
  const int zc0 = chunk * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zlo = dt.has_lo ? -1 : 0;
  const int zhi = dt.has_hi ? dt.nz : dt.nz - 1;
  const int zfirst = zc0 - 1, zlast = zc1;
  const bool ghost_lo = (zc0 == 0 && dt.has_lo), ghost_hi = (zc1 == dt.nz && dt.has_hi);   // this CTA reads ghost planes
  constexpr bool NEED_R = (MODE == STEP_JACOBI || MODE == STEP_PREC || MODE == STEP_UPDPREC);   // r's ghost planes (STEP_UPDPREC: posted
                                                                                                // by the CG initialisation, kept up to date locally afterwards)
  constexpr bool NEED_Z = (MODE == STEP_Z || MODE == STEP_HORNER || MODE == STEP_UPDPREC);      // z, w_{k-1}, q
  if (a.peer && ghost_lo) {  // the neighbours fill our ghost planes directly: the plane BELOW is the first one of the march
    if (tid == 0) {
      if (NEED_R) peer_spin(a.peer->hflag_from_prev, a.seq_halo, a.peer->error);
      if (NEED_Z) peer_spin(a.peer->zflag_from_prev, a.seq_zin, a.peer->error);
    }
    __syncthreads();
  }
  // where this rank's boundary planes of the output go in the neighbours' slabs (nullptr: nothing to export)
  T *exp_prev = nullptr, *exp_next = nullptr;
  if (a.peer) {
    if (POLY) {
      const int wi = cg_widx(a.w_out_scr, it, MODE == STEP_UPDPREC);
      exp_prev = dt.has_lo ? (T *)a.peer->wghost_at_prev[wi] : nullptr;
      exp_next = dt.has_hi ? (T *)a.peer->wghost_at_next[wi] : nullptr;
    } else if (MODE == STEP_INIT) {        // r's ghost planes at the neighbours
      exp_prev = dt.has_lo ? (T *)a.peer->rghost_at_prev : nullptr;
      exp_next = dt.has_hi ? (T *)a.peer->rghost_at_next : nullptr;
    } else if (MODE == STEP_Z && a.r2) {   // fused update: the neighbours form r - alpha q on their ghost planes
      exp_prev = dt.has_lo ? (T *)a.peer->qghost_at_prev : nullptr;
      exp_next = dt.has_hi ? (T *)a.peer->qghost_at_next : nullptr;
    }
  }
  bool stored_peer = false;
  // ---- TMA staging (Cfg::TMAD >= 2): per-warp ring [stage][array][row][64] and one mbarrier per stage
  constexpr int TMAD = Cfg::TMAD;
  static_assert(TMAD == 0 || MODE == STEP_INIT, "TMA staging prefetches planes ahead of the deferred peer-flag wait: STEP_INIT only");
  constexpr int NARR = Cfg::narr(MODE);
  constexpr int CSLOT = NARR - 1;                       // slot of the third array (p_old / q) when there is one
  MVTV_DYN_SMEM(smem_raw);
  T *ring = reinterpret_cast<T *>(smem_raw) + (size_t)warp * (TMAD > 0 ? TMAD : 1) * NARR * NR * 64;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(reinterpret_cast<T *>(smem_raw) + (size_t)Cfg::WARPS * TMAD * NARR * NR * 64) + warp * TMAD;
  const int rowlen = min(64, m0 - xw);                  // vertices of the strip inside the mesh (even)
  const int xl = xo - xw;                               // this lane's pair inside the staged row
  if (TMAD > 0) {
    if (lane == 0) {
      for (int s_ = 0; s_ < TMAD; ++s_) mbar_init(&mbar[s_], 1);
      mbar_init_fence();
    }
    __syncwarp();
  }
  // NOC: boundary class bits of this lane's outputs within a plane (bit 0: axis 0, bit 1: axis 1)
  int cls[RY][2];
#pragma unroll
  for (int j = 0; j < RY; ++j)
#pragma unroll
    for (int k = 0; k < 2; ++k)
      cls[j][k] = ((x + k == 0 || x + k == m0 - 1) ? 1 : 0) | ((y0 + j == 0 || y0 + j == m1 - 1) ? 2 : 0);

  // raw inputs of the next plane, in flight while the current one is consumed
  // rcc / rdd: own-row extras -- diag(c) (!NOC) ; dinv and r (STEP_HORNER) ; theta and p (STEP_UPDPREC)
  T ra[NR][2], rb_[NR][2], rc[NR][2], rcc[RY][2], rdd[RY][2], re1[RY][2], re2[RY][2], re3[RY][2], ha[NR], hb[NR], hc[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    ra[r][0] = ra[r][1] = rb_[r][0] = rb_[r][1] = rc[r][0] = rc[r][1] = T(0);
    ha[r] = hb[r] = hc[r] = T(0);
  }
#pragma unroll
  for (int j = 0; j < RY; ++j) rcc[j][0] = rcc[j][1] = rdd[j][0] = rdd[j][1] = re1[j][0] = re1[j][1] = re2[j][0] = re2[j][1] = re3[j][0] = re3[j][1] = T(0);
  auto load_plane = [&](int zz) {
    if (zz > zlast) return;
    if (a.peer && ghost_hi && zz == dt.nz) {   // the ghost plane ABOVE is the last one of the march: its flag is only needed now,
      if (lane == 0) {                         // a late neighbour is hidden behind the whole chunk (each warp waits for itself)
        if (NEED_R) peer_spin(a.peer->hflag_from_next, a.seq_halo, a.peer->error);
        if (NEED_Z) peer_spin(a.peer->zflag_from_next, a.seq_zin, a.peer->error);
      }
      __syncwarp();
    }
    const int zs = min(max(zz, zlo), zhi);
    const long long pb = (long long)(zs + 1) * dt.plane;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (TMAD == 0) {
        ld2(rr + pb + rowoff[r], ra[r]);
        if (STAGE_B) ld2(dinv + pb + rowoff[r], rb_[r]);
        if (stage_c) ld2(p_in + pb + rowoff[r], rc[r]);
      }
      if (edge) {
        ha[r] = rr[pb + rowoff_h[r]];
        if (STAGE_B) hb[r] = dinv[pb + rowoff_h[r]];
        if (stage_c) hc[r] = p_in[pb + rowoff_h[r]];
      }
    }
    if (zz >= zc0 && zz < zc1) {
#pragma unroll
      for (int j = 0; j < RY; ++j)
        if (valid[j]) {
          const long long ob = (long long)(zz + 1) * dt.plane + (long long)(y0 + j) * m0 + x;
          if (!NOC || MODE == STEP_INIT) ld2(a.c + ob, rcc[j]);
          if (MODE == STEP_HORNER) { ld2(dinv + ob, rcc[j]); ld2(rcur + ob, rdd[j]); }
          if (MODE == STEP_UPDPREC) { ld2(a.x + ob, rcc[j]); ld2(p_dir + ob, rdd[j]); }
          if (MODE == STEP_INIT) { ld2(dinv + ob, rdd[j]); ld2(a.oty + ob, re1[j]); ld2(a.v1 + ob, re2[j]); ld2(a.v2 + ob, re3[j]); }
        }
    }
  };

  // TMA staging: plane zz goes into ring stage (zz - zfirst) % TMAD; every lane issues at most one row copy per pass of the loop
  auto tma_issue = [&](int zz) {
    if (TMAD == 0 || zz > zlast) return;
    const int stg = (zz - zfirst) % (TMAD > 0 ? TMAD : 1);
    const int zs = min(max(zz, zlo), zhi);
    const long long pb = (long long)(zs + 1) * dt.plane + xw;
    const int narr_now = (NARR == 3 || (MODE == STEP_Z)) ? (stage_c ? NARR : NARR - 1) : NARR;   // no p_old in the first iteration
    if (lane == 0) mbar_expect_tx(&mbar[stg], (unsigned)(narr_now * NR * rowlen * sizeof(T)));
    __syncwarp();
    for (int c = lane; c < NARR * NR; c += 32) {
      const int ai = c / NR, r = c - ai * NR;
      if (ai >= narr_now) continue;
      const int yy = min(max(y0 - 1 + r, 0), m1 - 1);
      const T *src = (ai == 0) ? rr : ((ai == CSLOT && (NARR == 3 || MODE == STEP_Z)) ? p_in : dinv);
      bulk_g2s(ring + ((size_t)(stg * NARR + ai) * NR + r) * 64, src + pb + (long long)yy * m0, (unsigned)(rowlen * sizeof(T)), &mbar[stg]);
    }
  };
  auto tma_consume = [&](int zz) {
    if (TMAD == 0) return;
    const int k = zz - zfirst, stg = k % (TMAD > 0 ? TMAD : 1);
    mbar_wait(&mbar[stg], (unsigned)((k / (TMAD > 0 ? TMAD : 1)) & 1));
    const T *sp = ring + (size_t)stg * NARR * NR * 64 + xl;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      ld2(sp + (size_t)r * 64, ra[r]);
      if (STAGE_B) ld2(sp + (size_t)(NR + r) * 64, rb_[r]);
      if (stage_c && (NARR == 3 || MODE == STEP_Z)) ld2(sp + (size_t)(CSLOT * NR + r) * 64, rc[r]);
    }
  };

  T A0[RY][2], A1[RY][2], A2[RY][2], pcp[RY][2], cqp[RY][2], rcp[RY][2], dcp[RY][2];
#pragma unroll
  for (int j = 0; j < RY; ++j)
#pragma unroll
    for (int k = 0; k < 2; ++k) A0[j][k] = A1[j][k] = A2[j][k] = pcp[j][k] = cqp[j][k] = rcp[j][k] = dcp[j][k] = T(0);
  double red[3] = {0.0, 0.0, 0.0};   // [1]: r.r of STEP_UPDPREC / STEP_INIT, [2]: b.b of STEP_INIT

  if (TMAD > 0) {
#pragma unroll
    for (int d_ = 0; d_ < (TMAD > 0 ? TMAD - 1 : 0); ++d_) tma_issue(zfirst + d_);
  }
  load_plane(zfirst);
  for (int zz = zfirst; zz <= zlast; ++zz) {
    tma_issue(zz + TMAD - 1);   // into the stage the previous plane has just left
    tma_consume(zz);
    // ---- combine: p_new (z0, w) of plane zz on the RY+2 rows, at the pair and at the strip's halo element
    T v[NR][2], hv[NR], rown[RY][2], down[RY][2], cown[RY][2];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (MODE == STEP_UPDPREC) {
          ra[r][k] -= alpha * rc[r][k];                  // r_new = r - alpha q
          v[r][k] = rb_[r][k] * ra[r][k];
        } else {
          v[r][k] = STAGE_B ? rb_[r][k] * ra[r][k] : ra[r][k];
          if (!first) v[r][k] += beta * rc[r][k];
        }
      }
      if (MODE == STEP_UPDPREC) hv[r] = hb[r] * (ha[r] - alpha * hc[r]);
      else {
        hv[r] = STAGE_B ? hb[r] * ha[r] : ha[r];
        if (!first) hv[r] += beta * hc[r];
      }
      if (!xvalid) v[r][0] = v[r][1];                    // replicate vertex m0-1
    }
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        rown[j][k] = (MODE == STEP_HORNER) ? rdd[j][k]
                     : ((MODE == STEP_INIT) ? re1[j][k] + (T)a.rho * (re2[j][k] + (T)a.uscale * re3[j][k]) : ra[j + 1][k]);   // STEP_INIT: b
        down[j][k] = (MODE == STEP_HORNER) ? rcc[j][k] : ((MODE == STEP_INIT) ? rdd[j][k] : rb_[j + 1][k]);
        cown[j][k] = rcc[j][k];
      }
    if (MODE == STEP_UPDPREC && zz >= zc0 && zz < zc1) {   // theta and r_new of the own rows of an own plane
#pragma unroll
      for (int j = 0; j < RY; ++j)
        if (valid[j]) {
          const long long ob = (long long)(zz + 1) * dt.plane + (long long)(y0 + j) * m0 + x;
          st2(r_out + ob, ra[j + 1][0], ra[j + 1][1]);
          st2(a.x + ob, rcc[j][0] + alpha * rdd[j][0], rcc[j][1] + alpha * rdd[j][1]);
          red[1] += (double)ra[j + 1][0] * (double)ra[j + 1][0] + (double)ra[j + 1][1] * (double)ra[j + 1][1];
        }
    }
    if (MODE == STEP_UPDPREC && ((zz < 0 && ghost_lo) || (zz >= dt.nz && ghost_hi))) {
      // several GPUs: r_new on the ghost planes, the same arithmetic as on the rank that owns them -- r needs no exchange
#pragma unroll
      for (int j = 0; j < RY; ++j)
        if (valid[j]) st2(r_out + (long long)(zz + 1) * dt.plane + (long long)(y0 + j) * m0 + x, ra[j + 1][0], ra[j + 1][1]);
    }
    load_plane(zz + 1);                                  // the registers are free: the next plane goes in flight
    {
      const int zs = min(max(zz, zlo), zhi);
      const bool own = (zz == zs) && ((zz >= zc0 && zz < zc1) || (zz < 0 && zc0 == 0) || (zz >= dt.nz && zc1 == dt.nz));
      if (!POLY && own) {
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
          if (MODE == STEP_INIT) {     // r = b - (c theta + rhoM K theta)
            const T rv = rcp[j][k] - (cqp[j][k] * pv + rhoM * A0[j][k]);
            outv[k] = rv;
            red[0] += (double)rv * (double)(rv * dcp[j][k]);
            red[1] += (double)rv * (double)rv;
            red[2] += (double)rcp[j][k] * (double)rcp[j][k];
          } else if (MODE == STEP_HORNER) {   // w_k = D^-1 M w_{k-1} + pc0 z0,  D^-1 M w = w + dinv*rhoM*(K w - diag(K) w)
            const T dk = (T)st.diagK[cls[j][k] | bz];
            const T zv = pv + rhoM * dcp[j][k] * (A0[j][k] - dk * pv) + (T)a.pc0 * (dcp[j][k] * rcp[j][k]);
            outv[k] = zv;
            red[0] += (double)rcp[j][k] * (double)zv;
          } else if (POLY) {
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
        st2((POLY ? zo : (MODE == STEP_INIT ? a.r : a.q)) + ob, outv[0], outv[1]);
        if (MODE == STEP_INIT) st2(a.xold + ob, pcp[j][0], pcp[j][1]);
        {  // several GPUs: the boundary planes also go straight into the neighbours' ghost planes
          const long long q = (long long)(y0 + j) * m0 + x;
          if (exp_prev && zz - 1 == 0) { st2(exp_prev + q, outv[0], outv[1]); stored_peer = true; }
          if (exp_next && zz - 1 == dt.nz - 1) { st2(exp_next + q, outv[0], outv[1]); stored_peer = true; }
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
  if (stored_peer) __threadfence_system();   // one fence per thread, before the CTA reports to the grid reduction
  double *S = a.S, *raw = a.raw;
  const PeerTab *peer = a.peer;
  const unsigned long long sr = a.seq_red, szo = a.seq_zout;
  const int fold = a.fold, fin = a.final_pass;
  const bool post_z = peer && (POLY || (MODE == STEP_Z && a.r2));   // this launch exported boundary planes
  // the epilogue runs in ONE thread after every CTA of the grid has finished (and fenced) its stores
  auto signal_z = [peer, szo, post_z]() {
    if (!post_z) return;
    __threadfence_system();
    if (peer->has_lo) st_flag_sys(peer->zflag_at_prev, szo);
    if (peer->has_hi) st_flag_sys(peer->zflag_at_next, szo);
  };
  if (MODE == STEP_INIT) {      // {r.z, r.r, b.b}; r's ghost planes are posted on the halo flags
    const unsigned long long sh = a.seq_halo;
    grid_reduce<3, 3>(red, rb, [S, raw, peer, sr, sh, fold](const double (&res)[3]) {
      if (peer) {
        __threadfence_system();
        if (peer->has_lo) st_flag_sys(peer->hflag_at_prev, sh);
        if (peer->has_hi) st_flag_sys(peer->hflag_at_next, sh);
        peer_post(*peer, sr, res, 3);
        if (fold) {   // what k_cg_peer_commit_init does
          double v[3];
          peer_wait_sum(*peer, sr, v, 3);
          cg_commit_init(S, v);
        }
      } else if (raw) { raw[0] = res[0]; raw[1] = res[1]; raw[2] = res[2]; }
      else cg_commit_init(S, res);
    });
    return;
  }
  if (MODE == STEP_UPDPREC) {   // {r.z, r.r} of the next parity, the iteration count advances (several GPUs: folded commits only)
    double r2[2] = {red[0], red[1]};
    const int defer = a.defer_rr;
    grid_reduce<2, 2>(r2, rb, [S, peer, sr, fin, defer, signal_z](const double (&res)[2]) {
      double v[2] = {res[0], res[1]};
      if (peer) {
        signal_z();
        peer_post(*peer, sr, res, 2);
        if (!fin && defer) {   // r.r is collected by the last Horner pass; until then the slot must not read as converged
          v[1] = 1.0e300;
          cg_commit_update_prec(S, v + 1);
          return;
        }
        peer_wait_sum(*peer, sr, v, 2);
      }
      if (fin) cg_commit_update(S, v);
      else cg_commit_update_prec(S, v + 1);   // r.z comes from the last Horner pass
    });
    return;
  }
  double r1[1] = {red[0]};
  if (MODE == STEP_HORNER || MODE == STEP_PREC) {
    if (!fin && !peer) return;   // one GPU, not the last pass: nothing to reduce, nobody to signal
    const unsigned long long srr = a.seq_rr_pending;
    grid_reduce<1, 1>(r1, rb, [S, raw, peer, sr, srr, fold, fin, signal_z](const double (&res)[1]) {
      if (peer) {
        signal_z();
        if (!fin) return;
        peer_post(*peer, sr, res, 1);
        if (fold) {   // what k_cg_peer_commit_rz does
          double v[1];
          peer_wait_sum(*peer, sr, v, 1);
          cg_commit_rz(S, v);
          if (srr) {  // the fused update's r.r of this iteration (posted, not yet summed)
            double w[2];
            peer_wait_sum(*peer, srr, w, 2);
            S[2 * (((int)S[CS_ITERS]) & 1) + 1] = w[1];
          }
        }
      } else if (raw) raw[0] = res[0];
      else cg_commit_rz(S, res);
    });
    return;
  }
  grid_reduce<1, 1>(r1, rb, [S, raw, peer, sr, fold, signal_z](const double (&res)[1]) {   // STEP_JACOBI / STEP_Z: p.q
    if (peer) {
      signal_z();
      peer_post(*peer, sr, res, 1);
      if (fold) {   // what k_cg_peer_commit_pq does
        double v[1];
        peer_wait_sum(*peer, sr, v, 1);
        S[CS_PQ] = v[0];
      }
    } else if (raw) raw[0] = res[0];
    else S[CS_PQ] = res[0];
  });
}

}  // namespace mvtv
