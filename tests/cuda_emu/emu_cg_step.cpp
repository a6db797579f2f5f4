// Runs the CG step kernels' SOURCE (csrc/kernels.cuh, cg_step2d.cuh, cg_step3d.cuh) on the CPU SIMT emulator
// (cuda_emu.h) and checks, on small awkward meshes, that
//   * the shared-memory kernel k_cg_step, the shuffle kernels k_cg_step2d / k_cg_step3d and a plain host loop over the
//     clamped 3^P-point stencil agree on p_new, q = M p_new (or z = P r) and the reduced scalar;
//   * chunking of the marching axis does not change the result.
// k_cg_step and k_cg_step2d are validated on B200 against the oracle, so their agreement with the host loop validates the
// emulator; k_cg_step3d has not run on a GPU yet: this is its logic check.
//   usage: emu_cg_step            (exit code 0 = all checks passed; EMU_NEGATIVE=1 perturbs the kernels' rhoM by 1e-6 and must
//                                  make the comparisons fail: the self-test of the comparisons)
#include <cstdio>
#include <random>
#include <vector>

#include "cuda_emu.h"
// clang-format off
#include "../../multivartv_b200/csrc/kernels.cuh"
#include "../../multivartv_b200/csrc/zu_march.cuh"
#include "../../multivartv_b200/csrc/cg_step2d.cuh"
#include "../../multivartv_b200/csrc/cg_step3d.cuh"
#include "../../multivartv_b200/csrc/cg_fused2d.cuh"
#include "../../multivartv_b200/csrc/cg_init2d.cuh"
// clang-format on

using namespace mvtv;

struct Mesh {
  DimTab dt{};
  StencilTab st{};
  int P;
};

// the tables mvtv_plan::build_tables makes (solver.cu), single rank, reference variant
static Mesh make_mesh(std::vector<long long> m, const std::vector<double> &deltas) {
  Mesh M;
  const int P = (int)m.size();
  M.P = P;
  DimTab &dt = M.dt;
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
  std::vector<int> mask(K);
  std::vector<double> scale(K);
  for (int b = 0; b < K; ++b) {
    const int num = (b == 0) ? K : b;
    int Sm = 0;
    for (int a = 0; a < P; ++a)
      if ((num >> (P - 1 - a)) & 1) Sm |= 1 << a;
    double sc = 1.0;
    if (b != 0 && !deltas.empty())
      for (int a = 0; a < P; ++a)
        if (!((Sm >> a) & 1)) sc *= deltas[a];
    mask[b] = zu_block_mask(P, ZV_REFERENCE, b);
    scale[b] = sc;
  }
  StencilTab &st = M.st;
  st.npts = 1;
  for (int a = 0; a < P; ++a) st.npts *= 3;
  const double t3[3] = {-1.0, 2.0, -1.0};
  for (int o = 0; o < st.npts; ++o) {
    double c = 0.0;
    for (int b = 0; b < K; ++b) {
      double w = scale[b] * scale[b];
      int rem = o;
      for (int a = 0; a < P; ++a) {
        const int dgt = rem % 3;
        rem /= 3;
        if ((mask[b] >> a) & 1) w *= t3[dgt];
        else w *= (dgt == 1) ? 1.0 : 0.0;
      }
      c += w;
    }
    st.coef[o] = c;
  }
  for (int cls = 0; cls < (1 << P); ++cls) {
    double dsum = 0.0;
    for (int b = 0; b < K; ++b) {
      double w = scale[b] * scale[b];
      for (int a = 0; a < P; ++a)
        if ((mask[b] >> a) & 1) w *= ((cls >> a) & 1) ? (m[a] >= 2 ? 1.0 : 0.0) : 2.0;
      dsum += w;
    }
    st.diagK[cls] = dsum;
  }
  return M;
}

struct Problem {
  Mesh M;
  std::vector<double> r, dinv, p_in, c, z, x;
  double rhoM = 0.7, pc0 = 1.3, pc1 = -0.21, beta = 0.5;
};

static Problem make_problem(std::vector<long long> m, const std::vector<double> &deltas, unsigned seed) {
  Problem pb;
  pb.M = make_mesh(m, deltas);
  const DimTab &dt = pb.M.dt;
  std::mt19937_64 g(seed);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_int_distribution<int> cnt(0, 3);
  auto rnd = [&](std::vector<double> &v) {
    v.assign((size_t)dt.usz, 0.0);
    for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) v[(size_t)i] = nd(g);
  };
  rnd(pb.r); rnd(pb.p_in); rnd(pb.z); rnd(pb.x);
  pb.c.assign((size_t)dt.usz, 0.0);
  pb.dinv.assign((size_t)dt.usz, 0.0);
  for (long long zl = 0; zl < dt.nz; ++zl)
    for (long long q = 0; q < dt.plane; ++q) {
      int cls = 0;
      long long rem = q;
      for (int a = 0; a < pb.M.P - 1; ++a) {
        const long long ia = rem % dt.m[a];
        rem /= dt.m[a];
        cls |= ((ia == 0) || (ia + 1 == dt.m[a])) << a;
      }
      cls |= ((zl == 0) || (zl + 1 == dt.m[pb.M.P - 1])) << (pb.M.P - 1);
      const size_t i = (size_t)((zl + 1) * dt.plane + q);
      pb.c[i] = (double)cnt(g);
      pb.dinv[i] = 1.0 / (pb.c[i] + pb.rhoM * pb.M.st.diagK[cls]);   // k_make_dinv
    }
  return pb;
}

// host loop: (K v)[i] with clamped neighbour indices
static double stencil_at(const Problem &pb, const std::vector<double> &v, const std::vector<long long> &idx) {
  const DimTab &dt = pb.M.dt;
  const int P = pb.M.P;
  double acc = 0.0;
  for (int o = 0; o < pb.M.st.npts; ++o) {
    int rem = o;
    long long lin = dt.plane;   // ghost offset
    for (int a = 0; a < P; ++a) {
      const int d = rem % 3 - 1;
      rem /= 3;
      long long ia = idx[(size_t)a] + d;
      ia = std::min(std::max(ia, 0ll), dt.m[a] - 1);
      lin += ia * dt.stride[a];
    }
    acc += pb.M.st.coef[o] * v[(size_t)lin];
  }
  return acc;
}

struct Out {
  std::vector<double> p_out, q, z;
  double scalar = 0.0;
};

static void host_reference(const Problem &pb, int mode, Out &o) {
  const DimTab &dt = pb.M.dt;
  const int P = pb.M.P;
  std::vector<double> pn((size_t)dt.usz, 0.0);
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) {
    double v = (mode == STEP_Z) ? pb.z[(size_t)i] : pb.dinv[(size_t)i] * pb.r[(size_t)i];
    if (mode != STEP_PREC) v += pb.beta * pb.p_in[(size_t)i];
    pn[(size_t)i] = v;
  }
  o.p_out = pn;
  o.q.assign((size_t)dt.usz, 0.0);
  o.z.assign((size_t)dt.usz, 0.0);
  o.scalar = 0.0;
  std::vector<long long> idx((size_t)P);
  for (long long li = 0; li < dt.Nloc; ++li) {
    long long rem = li;
    for (int a = 0; a < P; ++a) { idx[(size_t)a] = rem % dt.m[a]; rem /= dt.m[a]; }
    const size_t i = (size_t)(dt.plane + li);
    const double qv = pb.c[i] * pn[i] + pb.rhoM * stencil_at(pb, pn, idx);
    if (mode == STEP_PREC) {
      const double zv = pb.pc0 * pn[i] + pb.pc1 * (pb.dinv[i] * qv);
      o.z[i] = zv;
      o.scalar += pb.r[i] * zv;
    } else {
      o.q[i] = qv;
      o.scalar += pn[i] * qv;
    }
  }
}

template <typename Launch>
static void run_kernel(const Problem &pb, int mode, Launch launch, Out &o) {
  const DimTab &dt = pb.M.dt;
  std::vector<double> r = pb.r, dinv = pb.dinv, p_in = pb.p_in, c = pb.c, z = pb.z, x = pb.x;
  o.p_out.assign((size_t)dt.usz, 0.0);
  o.q.assign((size_t)dt.usz, 0.0);
  std::vector<double> S(CS_N, 0.0);
  S[CS_ITERS] = 1.0;            // not the first iteration: cur = 1, beta = S[2] / S[0]
  S[CS_RZ0] = 1.0 / pb.beta;
  S[CS_RZ1] = 1.0;
  S[CS_RR1] = 1.0;
  S[CS_BB] = 1.0;
  CgArgs<double> a{};
  a.x = x.data();
  a.r = r.data();
  a.q = o.q.data();
  a.pbuf[1] = p_in.data();
  a.pbuf[0] = o.p_out.data();
  a.c = c.data();
  a.dinv = dinv.data();
  a.S = S.data();
  a.raw = nullptr;
  a.peer = nullptr;
  a.rhoM = pb.rhoM + (getenv("EMU_NEGATIVE") ? 1e-6 : 0.0);   // EMU_NEGATIVE=1: every kernel must now DISAGREE with the host loop
  a.rtol2 = 1e-26;
  a.z = z.data();
  a.pc0 = pb.pc0;
  a.pc1 = pb.pc1;
  a.prec = 1;
  std::vector<double> partials(1 << 16, 0.0);
  unsigned counter = 0;
  launch(a, RedBuf{partials.data(), &counter});
  o.z = z;
  o.scalar = (mode == STEP_PREC) ? S[2 * 1] : S[CS_PQ];   // cg_commit_rz writes the slot of the current parity
}

static int g_fail = 0;
static void compare(const char *what, const Problem &pb, int mode, const Out &a, const Out &b, double tol) {
  const DimTab &dt = pb.M.dt;
  double e1 = 0, e2 = 0;
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) {
    if (mode != STEP_PREC) {
      e1 = std::max(e1, std::fabs(a.p_out[(size_t)i] - b.p_out[(size_t)i]));
      e2 = std::max(e2, std::fabs(a.q[(size_t)i] - b.q[(size_t)i]));
    } else {
      e2 = std::max(e2, std::fabs(a.z[(size_t)i] - b.z[(size_t)i]));
    }
  }
  const double e3 = std::fabs(a.scalar - b.scalar) / std::max(1.0, std::fabs(b.scalar));
  const bool ok = e1 <= tol && e2 <= tol && e3 <= tol;
  if (!ok) {
    ++g_fail;
    std::printf("FAIL %s mode=%d mesh=%lldx%lldx%lld: p %.2e out %.2e scalar %.2e (%g vs %g)\n", what, mode, dt.m[0], dt.m[1], dt.m[2], e1, e2, e3,
                a.scalar, b.scalar);
  }
}

template <int MODE>
static void check_3d(std::vector<long long> m, const std::vector<double> &deltas, unsigned seed) {
  Problem pb = make_problem(m, deltas, seed);
  const DimTab dt = pb.M.dt;
  const StencilTab st = pb.M.st;
  Out ref, smem, shfl;
  host_reference(pb, MODE, ref);
  using Old = StepCfg<2, 32, 1, 16, 4, 1, 4>;   // StepShape<3> of solver.cu
  for (int nchunk = 1; nchunk <= 2; ++nchunk) {
    const int zchunk = (dt.nz + nchunk - 1) / nchunk;
    const unsigned nch = (unsigned)((dt.nz + zchunk - 1) / zchunk);
    run_kernel(pb, MODE, [&](const CgArgs<double> &a, RedBuf rb) {
      const unsigned tiles = (unsigned)(((dt.m[0] + Old::TX - 1) / Old::TX) * ((dt.m[1] + Old::TY - 1) / Old::TY));
      cuda_emu::launch(dim3(tiles, nch, 1), dim3(Old::NT, 1, 1), sizeof(double) * Old::smem_elems<MODE>(),
                       [&] { k_cg_step<double, Old, MODE>(dt, st, a, rb, zchunk); });
    }, smem);
    compare("k_cg_step vs host loop", pb, MODE, smem, ref, 1e-12);
    auto one = [&](auto cfg, const char *name) {
      using C3 = decltype(cfg);
      run_kernel(pb, MODE, [&](const CgArgs<double> &a, RedBuf rb) {
        const unsigned tiles = (unsigned)(((dt.m[0] + C3::TX - 1) / C3::TX) * ((dt.m[1] + C3::TY - 1) / C3::TY));
        cuda_emu::launch(dim3(tiles, nch, 1), dim3(C3::NT, 1, 1), 0, [&] { k_cg_step3d<double, C3, MODE>(dt, st, a, rb, zchunk); });
      }, shfl);
      compare(name, pb, MODE, shfl, ref, 1e-12);
      compare(name, pb, MODE, shfl, smem, 1e-12);
    };
    one(Step3dCfg<4, 2>{}, "k_cg_step3d<4,2>");
    one(Step3dCfg<4, 3>{}, "k_cg_step3d<4,3>");
    one(Step3dCfg<4, 4>{}, "k_cg_step3d<4,4>");
    one(Step3dCfg<8, 1>{}, "k_cg_step3d<8,1>");
    one(Step3dCfg<4, 2, 0, false>{}, "k_cg_step3d<4,2,c>");
  }
}

template <int MODE>
static void check_2d(std::vector<long long> m, const std::vector<double> &deltas, unsigned seed) {
  Problem pb = make_problem(m, deltas, seed);
  const DimTab dt = pb.M.dt;
  const StencilTab st = pb.M.st;
  Out ref, smem, shfl;
  host_reference(pb, MODE, ref);
  using Old = StepCfg<1, 256, 2, 1, 1, 1, 3>;   // StepShape<2> of solver.cu
  for (int nchunk = 1; nchunk <= 2; ++nchunk) {
    const int zchunk = (dt.nz + nchunk - 1) / nchunk;
    const unsigned nch = (unsigned)((dt.nz + zchunk - 1) / zchunk);
    run_kernel(pb, MODE, [&](const CgArgs<double> &a, RedBuf rb) {
      const unsigned tiles = (unsigned)((dt.m[0] + Old::TX - 1) / Old::TX);
      cuda_emu::launch(dim3(tiles, nch, 1), dim3(Old::NT, 1, 1), sizeof(double) * Old::smem_elems<MODE>(),
                       [&] { k_cg_step<double, Old, MODE>(dt, st, a, rb, zchunk); });
    }, smem);
    compare("2-D k_cg_step vs host loop", pb, MODE, smem, ref, 1e-12);
    auto one = [&](auto cfg, const char *name) {
      using C2 = decltype(cfg);
      run_kernel(pb, MODE, [&](const CgArgs<double> &a, RedBuf rb) {
        const unsigned tiles = (unsigned)((dt.m[0] + C2::TX - 1) / C2::TX);
        cuda_emu::launch(dim3(tiles, nch, 1), dim3(C2::NT, 1, 1), 0, [&] { k_cg_step2d<double, C2, MODE>(dt, st, a, rb, zchunk); });
      }, shfl);
      compare(name, pb, MODE, shfl, ref, 1e-12);
    };
    one(Step2dCfg<4, 1, 4, 0>{}, "k_cg_step2d<4,1,4>");
    one(Step2dCfg<8, 1, 2, 4, true>{}, "k_cg_step2d<8,1,2,noc>");
    one(Step2dCfg<4, 2, 2, 0>{}, "k_cg_step2d<4,2,2>");
    one(Step2dCfg<8, 1, 2, 4, true, true, true>{}, "k_cg_step2d<8,1,2,noc,dksel,idx32>");
    one(Step2dCfg<8, 1, 2, 4, false, false, true>{}, "k_cg_step2d<8,1,2,c,idx32>");
  }
}


// ---- slab partition: every rank's kernel, fed with ghost planes of its inputs, must reproduce its part of the global result
// (NCCL-style path: a.raw receives the rank's partial sum; p_new is also written on the ghost planes the first / last chunk
// keep up to date redundantly, so p never needs a halo exchange)
template <int MODE, typename LaunchFor>
static void check_slabs(const char *name, std::vector<long long> m, const std::vector<double> &deltas, unsigned seed, int world,
                        LaunchFor launch_for) {
  Problem pb = make_problem(m, deltas, seed);
  const DimTab gdt = pb.M.dt;
  const int P = pb.M.P;
  Out ref;
  host_reference(pb, MODE, ref);
  double total = 0.0;
  double e_p = 0, e_o = 0, e_g = 0;
  const long long mz = gdt.m[P - 1], basez = mz / world, extra = mz % world;
  for (int rank = 0; rank < world; ++rank) {
    DimTab dt = gdt;
    dt.z0 = rank * basez + std::min<long long>(rank, extra);
    dt.nz = (int)(basez + (rank < extra ? 1 : 0));
    dt.has_lo = rank > 0;
    dt.has_hi = rank < world - 1;
    dt.Nloc = dt.plane * dt.nz;
    dt.usz = dt.plane * (dt.nz + 2);
    auto slab = [&](const std::vector<double> &g) {   // local ghosted copy: planes z0-1 .. z0+nz of the global vector
      std::vector<double> v((size_t)dt.usz, 0.0);
      for (long long zl = -1; zl <= dt.nz; ++zl) {
        const long long gz = dt.z0 + zl;
        if (gz < 0 || gz >= mz) continue;
        std::copy(g.begin() + (gz + 1) * gdt.plane, g.begin() + (gz + 2) * gdt.plane, v.begin() + (zl + 1) * dt.plane);
      }
      return v;
    };
    std::vector<double> r = slab(pb.r), dinv = slab(pb.dinv), p_in = slab(pb.p_in), c = slab(pb.c), z = slab(pb.z);
    std::vector<double> p_out((size_t)dt.usz, 0.0), q((size_t)dt.usz, 0.0), S(CS_N, 0.0), raw(8, 0.0);
    S[CS_ITERS] = 1.0; S[CS_RZ0] = 1.0 / pb.beta; S[CS_RZ1] = 1.0; S[CS_RR1] = 1.0; S[CS_BB] = 1.0;
    CgArgs<double> a{};
    a.r = r.data(); a.q = q.data(); a.pbuf[1] = p_in.data(); a.pbuf[0] = p_out.data(); a.c = c.data(); a.dinv = dinv.data();
    a.S = S.data(); a.raw = raw.data(); a.peer = nullptr; a.rhoM = pb.rhoM; a.rtol2 = 1e-26; a.z = z.data(); a.pc0 = pb.pc0; a.pc1 = pb.pc1;
    a.prec = 1;
    std::vector<double> partials(1 << 16, 0.0);
    unsigned counter = 0;
    launch_for(dt, pb.M.st, a, RedBuf{partials.data(), &counter});
    total += raw[0];
    for (long long zl = 0; zl < dt.nz; ++zl)
      for (long long qq = 0; qq < dt.plane; ++qq) {
        const size_t li = (size_t)((zl + 1) * dt.plane + qq), gi = (size_t)((dt.z0 + zl + 1) * gdt.plane + qq);
        if (MODE != STEP_PREC) {
          e_p = std::max(e_p, std::fabs(p_out[li] - ref.p_out[gi]));
          e_o = std::max(e_o, std::fabs(q[li] - ref.q[gi]));
        } else {
          e_o = std::max(e_o, std::fabs(z[li] - ref.z[gi]));
        }
      }
    if (MODE != STEP_PREC)   // ghost planes of p_new
      for (int side = 0; side < 2; ++side) {
        if ((side == 0 && !dt.has_lo) || (side == 1 && !dt.has_hi)) continue;
        const long long zl = side == 0 ? -1 : dt.nz;
        for (long long qq = 0; qq < dt.plane; ++qq)
          e_g = std::max(e_g, std::fabs(p_out[(size_t)((zl + 1) * dt.plane + qq)] - ref.p_out[(size_t)((dt.z0 + zl + 1) * gdt.plane + qq)]));
      }
  }
  const double e_s = std::fabs(total - ref.scalar) / std::max(1.0, std::fabs(ref.scalar));
  if (!(e_p <= 1e-12 && e_o <= 1e-12 && e_g <= 1e-12 && e_s <= 1e-12)) {
    ++g_fail;
    std::printf("FAIL slabs %s mode=%d world=%d: p %.2e out %.2e ghost-p %.2e scalar %.2e\n", name, MODE, world, e_p, e_o, e_g, e_s);
  }
}

template <int MODE>
static void check_slabs_all(unsigned seed) {
  const std::vector<double> none, d3 = {0.3, 0.5, 2.0};
  for (int world : {2, 3}) {
    check_slabs<MODE>("k_cg_step 3-D", {12, 10, 9}, d3, seed, world, [](const DimTab &dt, const StencilTab &st, const CgArgs<double> &a, RedBuf rb) {
      using Old = StepCfg<2, 32, 1, 16, 4, 1, 4>;
      const unsigned tiles = (unsigned)(((dt.m[0] + Old::TX - 1) / Old::TX) * ((dt.m[1] + Old::TY - 1) / Old::TY));
      const int zchunk = (dt.nz + 1) / 2;
      cuda_emu::launch(dim3(tiles, (unsigned)((dt.nz + zchunk - 1) / zchunk), 1), dim3(Old::NT, 1, 1), sizeof(double) * Old::smem_elems<MODE>(),
                       [&] { k_cg_step<double, Old, MODE>(dt, st, a, rb, zchunk); });
    });
    check_slabs<MODE>("k_cg_step3d<4,2>", {12, 10, 9}, d3, seed, world, [](const DimTab &dt, const StencilTab &st, const CgArgs<double> &a, RedBuf rb) {
      using C3 = Step3dCfg<4, 2>;
      const unsigned tiles = (unsigned)(((dt.m[0] + C3::TX - 1) / C3::TX) * ((dt.m[1] + C3::TY - 1) / C3::TY));
      const int zchunk = (dt.nz + 1) / 2;
      cuda_emu::launch(dim3(tiles, (unsigned)((dt.nz + zchunk - 1) / zchunk), 1), dim3(C3::NT, 1, 1), 0,
                       [&] { k_cg_step3d<double, C3, MODE>(dt, st, a, rb, zchunk); });
    });
    check_slabs<MODE>("k_cg_step3d<8,1>", {70, 7, 8}, none, seed + 1, world, [](const DimTab &dt, const StencilTab &st, const CgArgs<double> &a, RedBuf rb) {
      using C3 = Step3dCfg<8, 1>;
      const unsigned tiles = (unsigned)(((dt.m[0] + C3::TX - 1) / C3::TX) * ((dt.m[1] + C3::TY - 1) / C3::TY));
      cuda_emu::launch(dim3(tiles, 1, 1), dim3(C3::NT, 1, 1), 0, [&] { k_cg_step3d<double, C3, MODE>(dt, st, a, rb, dt.nz); });
    });
    check_slabs<MODE>("k_cg_step2d", {66, 11}, none, seed + 2, world, [](const DimTab &dt, const StencilTab &st, const CgArgs<double> &a, RedBuf rb) {
      using C2 = Step2dCfg<8, 1, 2, 4, true>;
      const unsigned tiles = (unsigned)((dt.m[0] + C2::TX - 1) / C2::TX);
      const int zchunk = (dt.nz + 1) / 2;
      cuda_emu::launch(dim3(tiles, (unsigned)((dt.nz + zchunk - 1) / zchunk), 1), dim3(C2::NT, 1, 1), 0,
                       [&] { k_cg_step2d<double, C2, MODE>(dt, st, a, rb, zchunk); });
    });
  }
}


// ---- k_cg_updprec2d (vector update fused with the preconditioner): against x + a p, r - a q, P(D^-1 M) D^-1 r_new on the host
template <typename CF>
static void check_fused(const char *name, std::vector<long long> m, const std::vector<double> &deltas, unsigned seed, int iters) {
  Problem pb = make_problem(m, deltas, seed);
  const DimTab dt = pb.M.dt;
  const StencilTab st = pb.M.st;
  std::mt19937_64 g(seed + 100);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> q((size_t)dt.usz, 0.0);
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) q[(size_t)i] = nd(g);
  const double rz = 0.9, pq = 1.7, alpha = rz / pq;
  // host reference
  Problem pr = pb;
  std::vector<double> x_ref = pb.x, r_ref = pb.r;
  double rr_ref = 0.0;
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) {
    x_ref[(size_t)i] = pb.x[(size_t)i] + alpha * pb.p_in[(size_t)i];
    r_ref[(size_t)i] = pb.r[(size_t)i] - alpha * q[(size_t)i];
    rr_ref += r_ref[(size_t)i] * r_ref[(size_t)i];
  }
  pr.r = r_ref;
  Out ref;
  host_reference(pr, STEP_PREC, ref);   // z = P r_new, scalar = r_new . z
  // kernel
  const int cur = iters & 1;
  std::vector<double> x = pb.x, p = pb.p_in, r0 = pb.r, r1 = pb.r, dinv = pb.dinv, z((size_t)dt.usz, 0.0), S(CS_N, 0.0);
  std::vector<double> &r_in = cur ? r1 : r0, &r_out = cur ? r0 : r1;
  std::fill(r_out.begin(), r_out.end(), -777.0);   // stale contents of the other buffer must not matter
  (void)r_in;
  S[CS_ITERS] = (double)iters;
  S[2 * cur] = rz; S[2 * cur + 1] = 1.0; S[2 * (cur ^ 1)] = 0.3; S[2 * (cur ^ 1) + 1] = 2.0;
  S[CS_PQ] = pq; S[CS_BB] = 1.0;
  CgArgs<double> a{};
  a.x = x.data(); a.q = q.data(); a.pbuf[cur ^ 1] = p.data(); a.pbuf[cur] = nullptr; a.dinv = dinv.data(); a.S = S.data();
  a.rhoM = pb.rhoM; a.rtol2 = 1e-26; a.z = z.data(); a.pc0 = pb.pc0; a.pc1 = pb.pc1; a.prec = 1;
  std::vector<double> partials(1 << 16, 0.0);
  unsigned counter = 0;
  for (int nchunk = 1; nchunk <= 2; ++nchunk) {
    std::vector<double> xs = x, r0s = r0, r1s = r1, zs = z, Ss = S;
    a.x = xs.data(); a.z = zs.data(); a.S = Ss.data();
    const int zchunk = (dt.nz + nchunk - 1) / nchunk;
    const unsigned tiles = (unsigned)((dt.m[0] + CF::TX - 1) / CF::TX);
    cuda_emu::launch(dim3(tiles, (unsigned)((dt.nz + zchunk - 1) / zchunk), 1), dim3(CF::NT, 1, 1), 0,
                     [&] { a.r = r0s.data(); a.r2 = r1s.data(); k_cg_updprec2d<double, CF>(dt, st, a, RedBuf{partials.data(), &counter}, zchunk); });
    const std::vector<double> &rn = cur ? r0s : r1s;
    double ex = 0, er = 0, ez = 0;
    for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) {
      ex = std::max(ex, std::fabs(xs[(size_t)i] - x_ref[(size_t)i]));
      er = std::max(er, std::fabs(rn[(size_t)i] - r_ref[(size_t)i]));
      ez = std::max(ez, std::fabs(zs[(size_t)i] - ref.z[(size_t)i]));
    }
    const int nxt = cur ^ 1;
    const double es = std::max(std::fabs(Ss[2 * nxt] - ref.scalar) / std::max(1.0, std::fabs(ref.scalar)),
                               std::fabs(Ss[2 * nxt + 1] - rr_ref) / std::max(1.0, rr_ref));
    if (!(ex <= 1e-13 && er <= 1e-13 && ez <= 1e-12 && es <= 1e-12 && Ss[CS_ITERS] == iters + 1.0)) {
      ++g_fail;
      std::printf("FAIL fused %s mesh=%lldx%lld iters=%d chunks=%d: x %.2e r %.2e z %.2e scalars %.2e iters %g\n", name, dt.m[0], dt.m[1], iters, nchunk,
                  ex, er, ez, es, Ss[CS_ITERS]);
    }
  }
}


// ---- k_cg_init2d (marching / shuffle form) against k_cg_init (gather form, GPU-validated) and the host loop
static void check_init2d(std::vector<long long> m, const std::vector<double> &deltas, unsigned seed) {
  Problem pb = make_problem(m, deltas, seed);
  const DimTab dt = pb.M.dt;
  const StencilTab st = pb.M.st;
  std::mt19937_64 g(seed + 7);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> oty((size_t)dt.usz, 0.0), v1 = oty, v2 = oty;
  for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) { oty[(size_t)i] = nd(g); v1[(size_t)i] = nd(g); v2[(size_t)i] = nd(g); }
  const double rho = 1.3, usc = 0.5;
  auto run = [&](int which, int nchunk, std::vector<double> &r, std::vector<double> &xold, std::vector<double> &S) {
    std::vector<double> x = pb.x, c = pb.c, dinv = pb.dinv;
    r.assign((size_t)dt.usz, 0.0);
    xold.assign((size_t)dt.usz, 0.0);
    S.assign(CS_N, -1.0);
    CgArgs<double> a{};
    a.x = x.data(); a.xold = xold.data(); a.r = r.data(); a.c = c.data(); a.dinv = dinv.data(); a.oty = oty.data(); a.v1 = v1.data(); a.v2 = v2.data();
    a.S = S.data(); a.rho = rho; a.uscale = usc; a.rhoM = pb.rhoM; a.rtol2 = 1e-26;
    std::vector<double> partials((size_t)(1 << 16) * 3, 0.0);
    unsigned counter = 0;
    if (which == 0) {
      cuda_emu::launch(dim3((unsigned)((dt.plane + 255) / 256), (unsigned)dt.nz, 1), dim3(256, 1, 1), 0,
                       [&] { k_cg_init<double, 2>(dt, st, a, RedBuf{partials.data(), &counter}); });
    } else {
      const int zchunk = (dt.nz + nchunk - 1) / nchunk;
      cuda_emu::launch(dim3((unsigned)((dt.m[0] + 255) / 256), (unsigned)((dt.nz + zchunk - 1) / zchunk), 1), dim3(128, 1, 1), 0,
                       [&] { k_cg_init2d<double, 4>(dt, st, a, RedBuf{partials.data(), &counter}, zchunk); });
    }
  };
  std::vector<double> r0, xo0, S0, r1, xo1, S1;
  run(0, 1, r0, xo0, S0);
  for (int nchunk = 1; nchunk <= 3; ++nchunk) {
    run(1, nchunk, r1, xo1, S1);
    double er = 0, ex = 0, es = 0;
    for (long long i = dt.plane; i < dt.plane + dt.Nloc; ++i) {
      er = std::max(er, std::fabs(r0[(size_t)i] - r1[(size_t)i]));
      ex = std::max(ex, std::fabs(xo0[(size_t)i] - xo1[(size_t)i]));
    }
    for (int k : {CS_RZ0, CS_RR0, CS_BB, CS_ITERS, CS_PQ}) es = std::max(es, std::fabs(S0[(size_t)k] - S1[(size_t)k]) / std::max(1.0, std::fabs(S0[(size_t)k])));
    if (!(er <= 1e-12 && ex == 0.0 && es <= 1e-12)) {
      ++g_fail;
      std::printf("FAIL init2d mesh=%lldx%lld chunks=%d: r %.2e xold %.2e scalars %.2e\n", dt.m[0], dt.m[1], nchunk, er, ex, es);
    }
  }
}

int main() {
  const std::vector<double> none, d2 = {0.5, 0.25}, d3 = {0.3, 0.5, 2.0};
  // 2-D: the GPU-validated pair first (this validates the emulator itself)
  for (auto m : std::vector<std::vector<long long>>{{66, 5}, {130, 9}, {2, 7}, {258, 4}}) {
    check_2d<STEP_JACOBI>(m, none, 1);
    check_2d<STEP_Z>(m, d2, 2);
    check_2d<STEP_PREC>(m, none, 3);
  }
  std::printf("2-D checks done, failures so far: %d\n", g_fail);
  // 3-D: widths that are / are not multiples of the 64-vertex strip, heights that are / are not multiples of the row tiles
  for (auto m : std::vector<std::vector<long long>>{{12, 12, 5}, {66, 5, 4}, {2, 9, 3}, {70, 19, 3}, {6, 1, 4}}) {
    check_3d<STEP_JACOBI>(m, none, 4);
    check_3d<STEP_Z>(m, d3, 5);
    check_3d<STEP_PREC>(m, none, 6);
    check_3d<STEP_PREC>(m, d3, 7);
  }
  for (auto m : std::vector<std::vector<long long>>{{66, 5}, {130, 9}, {2, 7}, {258, 4}, {64, 12}})
    for (int iters : {0, 1, 4}) {
      check_fused<Fused2dCfg<8, 0>>("<8>", m, none, 21 + (unsigned)iters, iters);
      check_fused<Fused2dCfg<4, 0>>("<4>", m, d2, 31 + (unsigned)iters, iters);
    }
  for (auto m : std::vector<std::vector<long long>>{{66, 5}, {130, 9}, {2, 7}, {258, 4}, {64, 12}}) {
    check_init2d(m, none, 41);
    check_init2d(m, d2, 42);
  }
  check_slabs_all<STEP_JACOBI>(11);
  check_slabs_all<STEP_Z>(12);
  check_slabs_all<STEP_PREC>(13);
  std::printf("emu_cg_step: %d failure(s)\n", g_fail);
  return g_fail ? 1 : 0;
}
