// Whole ADMM passes on the CPU SIMT emulator: the kernel pipeline of mvtv_solve (RCPP mode, polynomial preconditioner, fixed
// rho) driven by a small host loop -- z/u init, then per pass k_make_dinv, CG initialisation, the CG iterations, fused z/u
// update -- once with the GPU-validated kernels (k_zu_march default tile, k_cg_init, k_cg_step / k_cg_step2d, k_cg_update) and
// once with each opt-in combination (k_cg_init2d, k_cg_updprec2d with its out-of-place residual, k_cg_step3d, other z/u tiles).
// theta and u after the passes must agree: this checks what the per-kernel drivers cannot -- the parity-selected buffers and
// scalar slots across iterations, the first-iteration preconditioner call of the fused path, chunking choices interacting.
//   usage: emu_solve [quick]     (exit code 0 = all combinations agree with the baseline pipeline; the full run takes ~30 s, `quick` a reduced set)
#include <cstdio>
#include <functional>
#include <random>
#include <string>
#include <vector>

#include "cuda_emu.h"
// clang-format off
#include "../../multivartv_b200/csrc/kernels.cuh"
#include "../../multivartv_b200/csrc/zu_march.cuh"
#include "../../multivartv_b200/csrc/cg_step2d.cuh"
#include "../../multivartv_b200/csrc/cg_step3d.cuh"
#include "../../multivartv_b200/csrc/cg_step3dh.cuh"
#include "../../multivartv_b200/csrc/cg_fused2d.cuh"
#include "../../multivartv_b200/csrc/cg_init2d.cuh"
#include "../../multivartv_b200/csrc/cg_horner2d.cuh"
// clang-format on

using namespace mvtv;

static int pdep_int(int j, int S) {
  int e = 0, k = 0;
  for (int a = 0; a < 8; ++a)
    if ((S >> a) & 1) {
      if ((j >> k) & 1) e |= 1 << a;
      ++k;
    }
  return e;
}

struct Tabs {
  DimTab dt{};
  BlockTab bt{};
  StencilTab st{};
  double cheb_bmax = 1.0;
  int P = 0;
};

// mvtv_plan::build_tables (solver.cu), single rank, reference variant, unit block scales
static Tabs make_tabs(std::vector<long long> m) {
  Tabs T;
  const int P = (int)m.size();
  T.P = P;
  DimTab &dt = T.dt;
  dt.P = P;
  long long s = 1;
  for (int a = 0; a < MVTV_MAXP; ++a) {
    dt.m[a] = a < P ? m[a] : 1;
    dt.stride[a] = a < P ? s : 0;
    if (a < P) s *= m[a];
  }
  dt.plane = dt.stride[P - 1];
  dt.nz = (int)m[P - 1];
  dt.Nloc = dt.plane * dt.nz;
  dt.usz = dt.plane * (dt.nz + 2);
  const int K = (1 << P) - 1;
  T.bt.K = K;
  for (int b = 0; b < K; ++b) {
    const int Sp = zu_block_mask(P, ZV_REFERENCE, b);
    T.bt.mask[b] = Sp;
    T.bt.scale[b] = 1.0;
    T.bt.nsub[b] = 1 << __builtin_popcount(Sp);
    for (int j = 0; j < T.bt.nsub[b]; ++j) {
      const int e = pdep_int(j, Sp);
      T.bt.sub[b][j] = e;
      long long o = 0;
      for (int a = 0; a < P; ++a)
        if ((e >> a) & 1) o += dt.stride[a];
      T.bt.off[b][j] = o;
    }
  }
  StencilTab &st = T.st;
  st.npts = 1;
  for (int a = 0; a < P; ++a) st.npts *= 3;
  const double t3[3] = {-1.0, 2.0, -1.0};
  for (int o = 0; o < st.npts; ++o) {
    double c = 0.0;
    for (int b = 0; b < K; ++b) {
      double w = 1.0;
      int rem = o;
      for (int a = 0; a < P; ++a) {
        const int dgt = rem % 3;
        rem /= 3;
        w *= ((T.bt.mask[b] >> a) & 1) ? t3[dgt] : ((dgt == 1) ? 1.0 : 0.0);
      }
      c += w;
    }
    st.coef[o] = c;
  }
  for (int cls = 0; cls < (1 << P); ++cls) {
    double dsum = 0.0;
    for (int b = 0; b < K; ++b) {
      double w = 1.0;
      for (int a = 0; a < P; ++a)
        if ((T.bt.mask[b] >> a) & 1) w *= ((cls >> a) & 1) ? (m[a] >= 2 ? 1.0 : 0.0) : 2.0;
      dsum += w;
    }
    st.diagK[cls] = dsum;
  }
  // Gershgorin bound on spec(D^-1 M) as in build_tables
  for (int cls = 0; cls < (1 << P); ++cls) {
    bool possible = true;
    for (int a = 0; a < P; ++a)
      if (!((cls >> a) & 1) && m[a] < 3) possible = false;
    if (!possible) continue;
    std::vector<double> row((size_t)st.npts, 0.0);
    for (int o = 0; o < st.npts; ++o) {
      int rem = o, tgt = 0, mul = 1;
      for (int a = 0; a < P; ++a) {
        int dgt = rem % 3;
        rem /= 3;
        if ((cls >> a) & 1) {
          if (m[a] < 2) dgt = 1;
          else if (dgt == 0) dgt = 1;
        }
        tgt += dgt * mul;
        mul *= 3;
      }
      row[(size_t)tgt] += st.coef[o];
    }
    int centre = 0, mul = 1;
    for (int a = 0; a < P; ++a) { centre += mul; mul *= 3; }
    double rowabs = 0.0;
    for (double v : row) rowabs += std::fabs(v);
    if (row[(size_t)centre] > 0.0) T.cheb_bmax = std::max(T.cheb_bmax, rowabs / row[(size_t)centre]);
  }
  return T;
}

// which kernels a pipeline uses
struct Pipeline {
  const char *name;
  int zu = 0;       // tile variant of k_zu_march (as MVTV_ZU_CFG)
  bool init2d = false, fused = false;
  int step = 0;     // 0: k_cg_step (shared memory), 1: k_cg_step2d / k_cg_step3d, 2: k_cg_step3dh
  int step_cfg = 0;
  int horner = 0;   // 2-D: degree of the Horner-form polynomial preconditioner k_cg_horner2d (0: k_cg_step2d<STEP_PREC>, degree 1)
};

struct State {
  std::vector<double> theta, xold, v1, v2, oty, c, dinv, r, r2, q, z, p0, p1, u0, u1, S, zr;
  int ucur = 0;
};

static std::vector<double> g_partials((size_t)(1 << 16) * ZR_N, 0.0);
static unsigned g_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};

static void launch_zu(const Tabs &T, const Pipeline &pl, State &s, double kappa, int init, bool with_prev) {
  const DimTab dt = T.dt;
  const BlockTab bt = T.bt;
  ZuArgs<double> a{};
  a.theta = s.theta.data();
  a.theta_prev = with_prev ? s.xold.data() : nullptr;
  a.u_old = (s.ucur ? s.u1 : s.u0).data();
  a.u_new = (s.ucur ? s.u0 : s.u1).data();
  a.v1 = s.v1.data();
  a.v2 = s.v2.data();
  a.kappa = kappa;
  a.uscale = 1.0;
  a.mode = MVTV_MODE_RCPP;
  a.init = init;
  a.red_out = s.zr.data();
  RedBuf rb{g_partials.data(), &g_counters[0]};
  auto go = [&](auto cfg) {
    using C = decltype(cfg);
    constexpr int Q = C::Q;
    const long long m0 = dt.m[0], m1 = Q >= 2 ? dt.m[1] : 1, m2 = Q >= 3 ? dt.m[2] : 1;
    const unsigned tiles = (unsigned)(((m0 + C::OX - 1) / C::OX) * ((m1 + C::OY - 1) / C::OY) * ((m2 + C::OW - 1) / C::OW));
    const int zchunk = (dt.nz + 1) / 2;
    cuda_emu::launch(dim3(tiles, (unsigned)((dt.nz + zchunk - 1) / zchunk), 1), dim3(C::NT, 1, 1), sizeof(double) * (size_t)C::SMEM_ELEMS,
                     [&] { k_zu_march<double, C, ZV_REFERENCE>(dt, bt, a, rb, zchunk); });
  };
  if (T.P == 2) go(ZuCfg<2, 256, 1, 1>{});
  else if (pl.zu == 1) go(ZuCfg<3, 32, 8, 1>{});
  else if (pl.zu == 3) go(ZuCfg<3, 64, 4, 1>{});
  else go(ZuCfg<3, 32, 16, 1>{});
}

static bool done(const State &s, double rtol2) {
  const int cur = ((int)s.S[CS_ITERS]) & 1;
  return s.S[2 * cur + 1] <= rtol2 * s.S[CS_BB];
}

static int cg_solve(const Tabs &T, const Pipeline &pl, State &s, double rho, double rhoM) {
  const DimTab dt = T.dt;
  const StencilTab st = T.st;
  const int P = T.P;
  cuda_emu::launch(dim3((unsigned)((dt.plane + 255) / 256), (unsigned)dt.nz + 2, 1), dim3(256, 1, 1), 0, [&] {
    if (P == 2) k_make_dinv<double, 2>(dt, st, s.c.data(), rhoM, s.dinv.data());
    else k_make_dinv<double, 3>(dt, st, s.c.data(), rhoM, s.dinv.data());
  });
  CgArgs<double> a{};
  a.x = s.theta.data(); a.xold = s.xold.data(); a.r = s.r.data(); a.q = s.q.data(); a.pbuf[0] = s.p0.data(); a.pbuf[1] = s.p1.data();
  a.c = s.c.data(); a.dinv = s.dinv.data(); a.oty = s.oty.data(); a.v1 = s.v1.data(); a.v2 = s.v2.data(); a.S = s.S.data();
  a.rho = rho; a.uscale = 1.0; a.rhoM = rhoM; a.rtol2 = 1e-24; a.z = s.z.data(); a.prec = 1;
  {
    const double b = T.cheb_bmax, lo = b / 30.0, th = 0.5 * (b + lo), de = 0.5 * (b - lo);
    const double T2 = 2.0 * (th / de) * (th / de) - 1.0;
    a.pc0 = 4.0 * th / (de * de * T2);
    a.pc1 = -2.0 / (de * de * T2);
  }
  const int zchunk = (dt.nz + 1) / 2;
  const unsigned nch = (unsigned)((dt.nz + zchunk - 1) / zchunk);
  // ---- initialisation
  if (pl.init2d) {
    cuda_emu::launch(dim3((unsigned)((dt.m[0] + 255) / 256), nch, 1), dim3(128, 1, 1), 0,
                     [&] { k_cg_init2d<double, 4>(dt, st, a, RedBuf{g_partials.data(), &g_counters[1]}, zchunk); });
  } else {
    cuda_emu::launch(dim3((unsigned)((dt.plane + 255) / 256), (unsigned)dt.nz, 1), dim3(256, 1, 1), 0, [&] {
      if (P == 2) k_cg_init<double, 2>(dt, st, a, RedBuf{g_partials.data(), &g_counters[1]});
      else k_cg_init<double, 3>(dt, st, a, RedBuf{g_partials.data(), &g_counters[1]});
    });
  }
  double hc[8] = {0};
  if (pl.horner) cheb_poly_coeffs(pl.horner, T.cheb_bmax, 30.0, hc);
  auto step = [&](int mode) {
    RedBuf rb{g_partials.data(), &g_counters[mode == STEP_PREC ? 5 : 2]};
    if (P == 2 && pl.horner && mode == STEP_PREC) {   // as mvtv_plan::cg_solve: one stencil pass per degree, ping-pong between q and z, ending in z
      const int d = pl.horner;
      const dim3 gh((unsigned)((dt.m[0] + 64 * 8 - 1) / (64 * 8)), nch, 1);
      for (int j = 1; j <= d; ++j) {
        double *w_out = ((d - j) % 2 == 0) ? s.z.data() : s.q.data();
        const double *w_in = ((d - j) % 2 == 0) ? s.q.data() : s.z.data();
        if (j == 1) cuda_emu::launch(gh, dim3(256, 1, 1), 0, [&] { k_cg_horner2d<double, 8, 4, true>(dt, st, a, nullptr, w_out, hc[d - 1], hc[d], j == d, rb, zchunk); });
        else cuda_emu::launch(gh, dim3(256, 1, 1), 0, [&] { k_cg_horner2d<double, 8, 4, false>(dt, st, a, w_in, w_out, hc[d - j], 0.0, j == d, rb, zchunk); });
      }
      return;
    }
    if (P == 2 && pl.step == 0) {
      using Old = StepCfg<1, 256, 2, 1, 1, 1, 3>;
      const unsigned tiles = (unsigned)((dt.m[0] + Old::TX - 1) / Old::TX);
      if (mode == STEP_PREC) cuda_emu::launch(dim3(tiles, nch, 1), dim3(Old::NT, 1, 1), sizeof(double) * Old::smem_elems<STEP_PREC>(), [&] { k_cg_step<double, Old, STEP_PREC>(dt, st, a, rb, zchunk); });
      else cuda_emu::launch(dim3(tiles, nch, 1), dim3(Old::NT, 1, 1), sizeof(double) * Old::smem_elems<STEP_Z>(), [&] { k_cg_step<double, Old, STEP_Z>(dt, st, a, rb, zchunk); });
    } else if (P == 2) {
      if (mode == STEP_PREC) {
        using C2 = Step2dCfg<8, 1, 2, 4, true>;
        cuda_emu::launch(dim3((unsigned)((dt.m[0] + C2::TX - 1) / C2::TX), nch, 1), dim3(C2::NT, 1, 1), 0, [&] { k_cg_step2d<double, C2, STEP_PREC>(dt, st, a, rb, zchunk); });
      } else {
        using C2 = Step2dCfg<4, 1, 4, 0>;
        cuda_emu::launch(dim3((unsigned)((dt.m[0] + C2::TX - 1) / C2::TX), nch, 1), dim3(C2::NT, 1, 1), 0, [&] { k_cg_step2d<double, C2, STEP_Z>(dt, st, a, rb, zchunk); });
      }
    } else if (pl.step == 0) {
      using Old = StepCfg<2, 32, 1, 16, 4, 1, 4>;
      const unsigned tiles = (unsigned)(((dt.m[0] + Old::TX - 1) / Old::TX) * ((dt.m[1] + Old::TY - 1) / Old::TY));
      if (mode == STEP_PREC) cuda_emu::launch(dim3(tiles, nch, 1), dim3(Old::NT, 1, 1), sizeof(double) * Old::smem_elems<STEP_PREC>(), [&] { k_cg_step<double, Old, STEP_PREC>(dt, st, a, rb, zchunk); });
      else cuda_emu::launch(dim3(tiles, nch, 1), dim3(Old::NT, 1, 1), sizeof(double) * Old::smem_elems<STEP_Z>(), [&] { k_cg_step<double, Old, STEP_Z>(dt, st, a, rb, zchunk); });
    } else if (pl.step == 2) {
      auto go = [&](auto cfg) {
        using CH = decltype(cfg);
        const unsigned tiles = (unsigned)(((dt.m[0] + CH::TX - 1) / CH::TX) * ((dt.m[1] + CH::TY - 1) / CH::TY));
        if (mode == STEP_PREC) cuda_emu::launch(dim3(tiles, nch, 1), dim3(CH::NT, 1, 1), 0, [&] { k_cg_step3dh<double, CH, STEP_PREC>(dt, st, a, rb, zchunk); });
        else cuda_emu::launch(dim3(tiles, nch, 1), dim3(CH::NT, 1, 1), 0, [&] { k_cg_step3dh<double, CH, STEP_Z>(dt, st, a, rb, zchunk); });
      };
      if (pl.step_cfg == 1) go(Step3dhCfg<6, 2>{});
      else go(Step3dhCfg<6, 1>{});
    } else {
      auto go = [&](auto cfg) {
        using C3 = decltype(cfg);
        const unsigned tiles = (unsigned)(((dt.m[0] + C3::TX - 1) / C3::TX) * ((dt.m[1] + C3::TY - 1) / C3::TY));
        if (mode == STEP_PREC) cuda_emu::launch(dim3(tiles, nch, 1), dim3(C3::NT, 1, 1), 0, [&] { k_cg_step3d<double, C3, STEP_PREC>(dt, st, a, rb, zchunk); });
        else cuda_emu::launch(dim3(tiles, nch, 1), dim3(C3::NT, 1, 1), 0, [&] { k_cg_step3d<double, C3, STEP_Z>(dt, st, a, rb, zchunk); });
      };
      if (pl.step_cfg == 1) go(Step3dCfg<4, 4>{});
      else if (pl.step_cfg == 6) go(Step3dCfg<8, 1>{});
      else go(Step3dCfg<4, 2>{});
    }
  };
  bool first_prec_done = false;
  int it = 0;
  for (; it < 400 && !done(s, a.rtol2); ++it) {
    if (!(pl.fused && first_prec_done)) {
      step(STEP_PREC);
      first_prec_done = true;
    }
    step(STEP_Z);
    if (pl.fused) {
      using CF = Fused2dCfg<8, 0>;
      cuda_emu::launch(dim3((unsigned)((dt.m[0] + CF::TX - 1) / CF::TX), nch, 1), dim3(CF::NT, 1, 1), 0, [&] {
        k_cg_updprec2d<double, CF>(dt, st, a, s.r.data(), s.r2.data(), RedBuf{g_partials.data(), &g_counters[3]}, zchunk);
      });
    } else {
      cuda_emu::launch(dim3(8, 1, 1), dim3(256, 1, 1), 0, [&] { k_cg_update<double>(a, dt.plane, dt.Nloc, RedBuf{g_partials.data(), &g_counters[3]}); });
    }
  }
  return it;
}

static int g_fail = 0;

static void run_pipeline(const Tabs &T, const Pipeline &pl, State &s, int passes, int *inner) {
  const double lambda = 0.8, rho = 0.4;
  *inner = 0;
  launch_zu(T, pl, s, 0.0, 1, false);           // alpha = D theta: D^T alpha, D^T u (solver.cu: launch_zu(0, uscale, mode, 1, false))
  for (int k = 0; k < passes; ++k) {
    *inner += cg_solve(T, pl, s, rho, rho);
    launch_zu(T, pl, s, lambda / rho, 0, true);
    s.ucur ^= 1;
  }
}

static int g_passes = 3;

static void check(std::vector<long long> m, unsigned seed, const std::vector<Pipeline> &pipes) {
  const Tabs T = make_tabs(m);
  const DimTab &dt = T.dt;
  std::mt19937_64 g(seed);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_int_distribution<int> cnt(0, 3);
  State init;
  const size_t n = (size_t)dt.usz;
  for (auto *v : {&init.theta, &init.xold, &init.v1, &init.v2, &init.oty, &init.c, &init.dinv, &init.r, &init.r2, &init.q, &init.z, &init.p0, &init.p1})
    v->assign(n, 0.0);
  init.u0.assign(n * (size_t)T.bt.K, 0.0);
  init.u1.assign(n * (size_t)T.bt.K, 0.0);
  init.S.assign(CS_N, 0.0);
  init.zr.assign(8, 0.0);
  double meany = 0.0;
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) {
    init.c[(size_t)i] = (double)cnt(g);
    init.oty[(size_t)i] = init.c[(size_t)i] * (1.0 + nd(g));
    meany += init.oty[(size_t)i];
  }
  meany /= std::max(1.0, (double)dt.Nloc);
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) init.theta[(size_t)i] = meany;
  std::fill(init.r2.begin(), init.r2.end(), -555.0);     // the second residual buffer starts with garbage
  State base = init;
  int inner0 = 0;
  run_pipeline(T, pipes[0], base, g_passes, &inner0);
  for (size_t k = 1; k < pipes.size(); ++k) {
    State s = init;
    int inner = 0;
    run_pipeline(T, pipes[k], s, g_passes, &inner);
    double et = 0, eu = 0;
    for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) et = std::max(et, std::fabs(s.theta[(size_t)i] - base.theta[(size_t)i]));
    const std::vector<double> &ua = s.ucur ? s.u1 : s.u0, &ub = base.ucur ? base.u1 : base.u0;
    for (int b = 0; b < T.bt.K; ++b)
      for (long long li = 0; li < dt.Nloc; ++li) {
        long long rem = li;
        bool exists = true;
        for (int ax = 0; ax < dt.P; ++ax) {
          const long long ia = rem % dt.m[ax];
          rem /= dt.m[ax];
          if (((T.bt.mask[b] >> ax) & 1) && ia + 1 >= dt.m[ax]) exists = false;
        }
        if (!exists) continue;
        const size_t i = (size_t)b * n + (size_t)(dt.plane + li);
        eu = std::max(eu, std::fabs(ua[i] - ub[i]));
      }
    // degree >= 2 preconditioners must need clearly fewer iterations than degree 1; everything else the same number
    const bool ok = et <= 1e-9 && eu <= 1e-9 && (pipes[k].horner >= 2 ? (inner < inner0 && 10 * inner >= 4 * inner0) : std::abs(inner - inner0) <= 3);
    std::printf("%s mesh=%lldx%lldx%lld %-28s: max|dtheta|=%.2e max|du|=%.2e CG iterations %d (baseline %d)\n", ok ? "ok  " : "FAIL", dt.m[0], dt.m[1],
                dt.m[2], pipes[k].name, et, eu, inner, inner0);
    if (!ok) ++g_fail;
  }
}

int main(int argc, char **argv) {
  if (argc > 1 && std::string(argv[1]) == "quick") {   // the default CPU test: two passes, one mesh per dimension
    g_passes = 2;
    check({66, 6}, 1, {{"baseline (k_cg_step2d)", 0, false, false, 1, 0}, {"k_cg_step2d + init2d + fused", 0, true, true, 1, 0}});
    check({12, 6, 4}, 3, {{"baseline (k_cg_step)", 0, false, false, 0, 0}, {"k_cg_step3d<4,2> + zu tile 1", 1, false, false, 1, 0},
                          {"k_cg_step3dh<6,2>", 0, false, false, 2, 1}});
    std::printf("emu_solve: %d failure(s)\n", g_fail);
    return g_fail ? 1 : 0;
  }
  const std::vector<Pipeline> p2 = {{"baseline (k_cg_step)", 0, false, false, 0, 0},
                                    {"k_cg_step2d", 0, false, false, 1, 0},
                                    {"k_cg_step2d + init2d", 0, true, false, 1, 0},
                                    {"k_cg_step2d + fused", 0, false, true, 1, 0},
                                    {"k_cg_step2d + init2d + fused", 0, true, true, 1, 0},
                                    {"k_cg_horner2d degree 1", 0, false, false, 1, 0, 1},
                                    {"k_cg_horner2d degree 2", 0, false, false, 1, 0, 2},
                                    {"k_cg_horner2d degree 3", 0, false, false, 1, 0, 3},
                                    {"k_cg_horner2d degree 4 + init2d", 0, true, false, 1, 0, 4}};
  check({66, 10}, 1, p2);
  check({130, 7}, 2, p2);
  const std::vector<Pipeline> p3 = {{"baseline (k_cg_step)", 0, false, false, 0, 0},
                                    {"k_cg_step3d<4,2>", 0, false, false, 1, 0},
                                    {"k_cg_step3d<4,4> + zu tile 1", 1, false, false, 1, 1},
                                    {"k_cg_step3d<8,1> + zu tile 3", 3, false, false, 1, 6},
                                    {"k_cg_step3dh<6,1>", 0, false, false, 2, 0},
                                    {"k_cg_step3dh<6,2> + zu tile 1", 1, false, false, 2, 1}};
  check({12, 10, 6}, 3, p3);
  check({66, 5, 5}, 4, p3);
  std::printf("emu_solve: %d failure(s)\n", g_fail);
  return g_fail ? 1 : 0;
}
