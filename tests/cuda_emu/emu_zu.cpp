// The fused z/u kernels on the CPU SIMT emulator: k_zu (gather form) and k_zu_march (scatter form, marching) in every tile
// variant solver.cu can select (MVTV_ZU_CFG) must agree on u_new, D^T alpha, D^T u_new and the six reductions, on meshes whose
// extents are / are not multiples of the tiles, single rank.  The two default configurations are validated on B200 against
// the oracle; the other tile shapes are checked here before they ever run on a GPU.
//   usage: emu_zu      (exit code 0 = all checks passed; EMU_NEGATIVE=1 perturbs kappa of the marching kernel: must fail)
#include <cstdio>
#include <random>
#include <vector>

#include "cuda_emu.h"
// clang-format off
#include "../../multivartv_b200/csrc/kernels.cuh"
#include "../../multivartv_b200/csrc/zu_march.cuh"
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
};
static Tabs make_tabs(std::vector<long long> m, const std::vector<double> &deltas) {
  Tabs T;
  const int P = (int)m.size();
  DimTab &dt = T.dt;
  dt.P = P;
  long long s = 1;
  for (int a = 0; a < MVTV_MAXP; ++a) {
    dt.m[a] = a < P ? m[a] : 1;
    dt.stride[a] = a < P ? s : 0;
    if (a < P) s *= m[a];
  }
  dt.plane = dt.stride[P - 1];
  dt.z0 = 0;
  dt.nz = (int)m[P - 1];
  dt.has_lo = dt.has_hi = 0;
  dt.Nloc = dt.plane * dt.nz;
  dt.usz = dt.plane * (dt.nz + 2);
  const int K = (1 << P) - 1;
  T.bt.K = K;
  for (int b = 0; b < K; ++b) {
    const int num = (b == 0) ? K : b;
    int Sm = 0;
    for (int a = 0; a < P; ++a)
      if ((num >> (P - 1 - a)) & 1) Sm |= 1 << a;
    double sc = 1.0;
    if (b != 0 && !deltas.empty())
      for (int a = 0; a < P; ++a)
        if (!((Sm >> a) & 1)) sc *= deltas[a];
    const int Sp = zu_block_mask(P, ZV_REFERENCE, b);
    T.bt.mask[b] = Sp;
    T.bt.scale[b] = sc;
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
  return T;
}

struct ZOut {
  std::vector<double> u_new, v1, v2, red;
};

static int g_fail = 0;

template <typename Launch>
static ZOut run(const Tabs &T, const std::vector<double> &theta, const std::vector<double> &theta_prev, const std::vector<double> &u_old,
                int mode, double kappa, Launch launch) {
  const DimTab &dt = T.dt;
  ZOut o;
  o.u_new.assign(u_old.size(), 0.0);
  o.v1.assign((size_t)dt.usz, 0.0);
  o.v2.assign((size_t)dt.usz, 0.0);
  o.red.assign(8, 0.0);
  ZuArgs<double> a{};
  a.theta = theta.data();
  a.theta_prev = theta_prev.data();
  a.u_old = u_old.data();
  a.u_new = o.u_new.data();
  a.v1 = o.v1.data();
  a.v2 = o.v2.data();
  a.kappa = kappa;
  a.uscale = 0.5;
  a.mode = mode;
  a.init = 0;
  a.red_out = o.red.data();
  std::vector<double> partials((size_t)(1 << 16) * ZR_N, 0.0);
  unsigned counter = 0;
  launch(a, RedBuf{partials.data(), &counter});
  return o;
}

static void compare(const char *name, const Tabs &T, const ZOut &a, const ZOut &ref) {
  const DimTab &dt = T.dt;
  double eu = 0, ev = 0, er = 0;
  // rows that exist: block b, vertex v with every axis of the block's set below the high boundary
  const ZOut &b = ref;
  for (int blk = 0; blk < T.bt.K; ++blk)
    for (long long li = 0; li < dt.Nloc; ++li) {
      long long rem = li;
      bool exists = true;
      for (int ax = 0; ax < dt.P; ++ax) {
        const long long ia = rem % dt.m[ax];
        rem /= dt.m[ax];
        if (((T.bt.mask[blk] >> ax) & 1) && ia + 1 >= dt.m[ax]) exists = false;
      }
      if (!exists) continue;
      const size_t i = (size_t)blk * (size_t)dt.usz + (size_t)(dt.plane + li);
      eu = std::max(eu, std::fabs(a.u_new[i] - b.u_new[i]));
    }
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) {
    ev = std::max(ev, std::fabs(a.v1[(size_t)i] - b.v1[(size_t)i]));
    ev = std::max(ev, std::fabs(a.v2[(size_t)i] - b.v2[(size_t)i]));
  }
  for (int k = 0; k < ZR_N; ++k) er = std::max(er, std::fabs(a.red[(size_t)k] - b.red[(size_t)k]) / std::max(1.0, std::fabs(b.red[(size_t)k])));
  if (!(eu <= 1e-13 && ev <= 1e-12 && er <= 1e-12)) {
    ++g_fail;
    std::printf("FAIL %s mesh=%lldx%lldx%lldx%lld: u %.2e D^T %.2e reductions %.2e\n", name, dt.m[0], dt.m[1], dt.m[2], dt.m[3], eu, ev, er);
  }
}

template <typename Cfg>
static ZOut march(const Tabs &T, const std::vector<double> &theta, const std::vector<double> &theta_prev, const std::vector<double> &u_old,
                  int mode, double kappa, int nchunk) {
  const DimTab dt = T.dt;
  const BlockTab bt = T.bt;
  constexpr int Q = Cfg::Q;
  const long long m0 = dt.m[0], m1 = Q >= 2 ? dt.m[1] : 1, m2 = Q >= 3 ? dt.m[2] : 1;
  const unsigned tiles = (unsigned)(((m0 + Cfg::OX - 1) / Cfg::OX) * ((m1 + Cfg::OY - 1) / Cfg::OY) * ((m2 + Cfg::OW - 1) / Cfg::OW));
  const int zchunk = (dt.nz + nchunk - 1) / nchunk;
  const unsigned nch = (unsigned)((dt.nz + zchunk - 1) / zchunk);
  const double kap = kappa + (getenv("EMU_NEGATIVE") ? 1e-6 : 0.0);
  return run(T, theta, theta_prev, u_old, mode, kap, [&](const ZuArgs<double> &a, RedBuf rb) {
    cuda_emu::launch(dim3(tiles, nch, 1), dim3(Cfg::NT, 1, 1), sizeof(double) * (size_t)Cfg::SMEM_ELEMS,
                     [&] { k_zu_march<double, Cfg, ZV_REFERENCE>(dt, bt, a, rb, zchunk); });
  });
}

static void check(std::vector<long long> m, const std::vector<double> &deltas, unsigned seed) {
  const Tabs T = make_tabs(m, deltas);
  const DimTab dt = T.dt;
  const BlockTab bt = T.bt;
  std::mt19937_64 g(seed);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> theta((size_t)dt.usz, 0.0), theta_prev((size_t)dt.usz, 0.0), u_old((size_t)dt.usz * (size_t)bt.K, 0.0);
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) { theta[(size_t)i] = nd(g); theta_prev[(size_t)i] = nd(g); }
  for (auto &v : u_old) v = 0.3 * nd(g);
  for (int mode : {MVTV_MODE_CPP, MVTV_MODE_RCPP}) {
    const double kappa = mode == MVTV_MODE_CPP ? 0.4 : 0.15;
    const ZOut ref = run(T, theta, theta_prev, u_old, mode, kappa, [&](const ZuArgs<double> &a, RedBuf rb) {
      cuda_emu::launch(dim3((unsigned)((dt.plane + 255) / 256), (unsigned)(dt.nz + dt.has_lo), 1), dim3(256, 1, 1), 0,
                       [&] { k_zu<double>(dt, bt, a, rb); });
    });
    for (int nchunk = 1; nchunk <= 2; ++nchunk) {
      if (dt.P == 2) {
        compare("ZuCfg<2,256,1,1>", T, march<ZuCfg<2, 256, 1, 1>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
      } else if (dt.P == 3) {
        compare("ZuCfg<3,32,16,1>", T, march<ZuCfg<3, 32, 16, 1>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<3,32,8,1>", T, march<ZuCfg<3, 32, 8, 1>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<3,64,8,1>", T, march<ZuCfg<3, 64, 8, 1>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<3,64,4,1>", T, march<ZuCfg<3, 64, 4, 1>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<3,16,16,1>", T, march<ZuCfg<3, 16, 16, 1>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
      } else {
        compare("ZuCfg<4,16,8,4>", T, march<ZuCfg<4, 16, 8, 4>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<4,16,4,4>", T, march<ZuCfg<4, 16, 4, 4>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<4,8,8,4>", T, march<ZuCfg<4, 8, 8, 4>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<4,8,8,8>", T, march<ZuCfg<4, 8, 8, 8>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
        compare("ZuCfg<4,16,8,2>", T, march<ZuCfg<4, 16, 8, 2>>(T, theta, theta_prev, u_old, mode, kappa, nchunk), ref);
      }
    }
  }
}

int main() {
  const std::vector<double> none;
  check({70, 9}, none, 1);
  check({300, 5}, {0.5, 0.25}, 2);
  check({33, 17, 5}, none, 3);
  check({12, 12, 4}, {0.3, 0.5, 2.0}, 4);
  check({70, 5, 3}, none, 5);
  check({6, 6, 6, 3}, none, 6);
  check({17, 9, 5, 4}, none, 7);
  check({5, 5, 5, 5}, {0.5, 0.25, 2.0, 3.0}, 8);
  std::printf("emu_zu: %d failure(s)\n", g_fail);
  return g_fail ? 1 : 0;
}
