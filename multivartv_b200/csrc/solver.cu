// Host driver + C ABI of libmvtv_b200.so: plan construction (block table of D, stencil of D^T D),
// the ADMM loop of admm_update (cpp-code/solvers.cpp:90-130, rcpp solvers.cpp:96-136,
// code/solvers.py:54-76) and the matrix-free PCG that replaces arma::spsolve (cpp-code/solvers.cpp:116).
// There is no CPU fallback anywhere in this file: without a usable CUDA device every entry point fails.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <vector>

#include "cg_step2d.cuh"
#include "cg_step3d.cuh"
#include "cg_fused2d.cuh"
#include "cg_init2d.cuh"
#include "cg_horner2d.cuh"
#include "kernels.cuh"
#include "setup.h"
#include "zu_march.cuh"

namespace mvtv {

static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (torch's bundled libnccl.so.2 is already mapped when the caller is a
// torch.distributed process; the single-GPU path never touches NCCL)
// ------------------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;

  void load() {
    if (handle) return;
    const char *names[] = {getenv("MVTV_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};   // MVTV_NCCL_LIB: a specific NCCL build
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) throw Error(MVTV_ERR_CUDA, std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
#define MVTV_SYM(field, name)                                                       \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                   \
  if (!field) throw Error(MVTV_ERR_CUDA, std::string("NCCL symbol missing: ") + name)
    MVTV_SYM(CommInitRank, "ncclCommInitRank");
    MVTV_SYM(CommDestroy, "ncclCommDestroy");
    MVTV_SYM(GetUniqueId, "ncclGetUniqueId");
    MVTV_SYM(AllReduce, "ncclAllReduce");
    MVTV_SYM(AllGather, "ncclAllGather");
    MVTV_SYM(Send, "ncclSend");
    MVTV_SYM(Recv, "ncclRecv");
    MVTV_SYM(GroupStart, "ncclGroupStart");
    MVTV_SYM(GroupEnd, "ncclGroupEnd");
    MVTV_SYM(GetErrorString, "ncclGetErrorString");
#undef MVTV_SYM
  }
};
static NcclApi g_nccl;

#define MVTV_NCCL(expr)                                                                             \
  do {                                                                                              \
    ncclResult_t _r = (expr);                                                                       \
    if (_r != ncclSuccess)                                                                          \
      throw ::mvtv::Error(MVTV_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));    \
  } while (0)

static inline int pdep_int(int j, int S) {
  int e = 0, k = 0;
  for (int a = 0; a < 8; ++a)
    if ((S >> a) & 1) {
      if ((j >> k) & 1) e |= 1 << a;
      ++k;
    }
  return e;
}

template <typename T> struct NcclType;
template <> struct NcclType<double> { static constexpr ncclDataType_t v = ncclFloat64; };
template <> struct NcclType<float> { static constexpr ncclDataType_t v = ncclFloat32; };

}  // namespace mvtv

using namespace mvtv;

enum { CGF_RING = 0, CGF_STRIP2D = 1, CGF_STRIP3D = 2 };

// ------------------------------------------------------------------------------------------------
struct mvtv_plan {
  int p = 0;            // user-visible number of axes
  int dtype = MVTV_F64;
  int variant = 0;
  int device = 0;
  int rank = 0, world = 1;
  DimTab dt{};
  BlockTab bt{};
  RowTab rt{};
  StencilTab st{};
  long long N = 0;      // global vertices
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ncclComm_t comm = nullptr;

  // device state (element type = dtype)
  void *theta = nullptr, *xold = nullptr, *v1 = nullptr, *v2 = nullptr, *oty = nullptr, *cnt = nullptr;
  void *r = nullptr, *q = nullptr, *dinv = nullptr, *zbuf = nullptr;
  void *pbuf[2] = {nullptr, nullptr};
  double dinv_rho = NAN;  // rhoM the inverse diagonal was built for
  void *u[2] = {nullptr, nullptr};
  int ucur = 0;
  double uscale = 1.0;
  double rho_state = 0.0;
  bool have_u_state = false, have_theta_state = false;

  double *S = nullptr, *zr = nullptr, *raw = nullptr;   // device scalars
  double *partials = nullptr;
  unsigned *counters = nullptr;  // [0] zu, [1] cg_init, [2] spmv, [3] update, [4] misc
  double *h_scal = nullptr;      // pinned host mirror (16 doubles)
  dim3 grid, block;
  unsigned nblocks = 0;

  // points
  long long n = 0;
  long long *vid = nullptr;
  double mean_y = 0.0;
  bool have_points = false;

  double *staging = nullptr;
  size_t staging_bytes = 0;
  // grow-only scratch of set_points (staged host inputs; sort keys + CUB temp) and capacity of vid: a plan that is
  // re-binned (CV folds, repeated end-to-end calls) never goes back to cudaMalloc / cudaFree
  void *in_buf = nullptr, *sort_buf = nullptr;
  size_t in_bytes = 0, sort_bytes = 0;
  long long vid_cap = 0;
  long long launches = 0;
  int last_cg_iters = 8;
  int last_cg_prec = 0;     // polynomial degree of the previous x-update (0 = Jacobi)
  double jac_estimate = 100.0; // MVTV_PRECOND_AUTO: decaying maximum of the Jacobi-equivalent iteration counts (a cold start is
                               // assumed hard: the first x-update from theta = mean(y) is the longest of a solve)
  // peer-memory collectives of the CG loop (CUDA IPC); falls back to NCCL when unavailable or MVTV_COMM=nccl
  unsigned char *cb = nullptr;          // this rank's comm buffer (slots, flags, halo flags, error word)
  PeerTab *d_peer = nullptr;            // device copy of the peer table, nullptr = NCCL path
  bool fold_commit = true;              // peer path: the reducing kernel's last thread commits the CG scalars (MVTV_FOLD_COMMIT=0: separate commit launches)
  std::vector<void *> ipc_opened;
  unsigned long long red_seq = 0, halo_seq = 0, zhalo_seq = 0;
  double cheb_bmax = 0.0;   // bound on the spectrum of D^-1 (diag(c) + s D^T D), independent of s and c
  double cheb_kappa = 30.0; // the polynomial preconditioner is the Chebyshev one on [bmax/kappa, bmax]
  // which kernels run the stencil passes of the CG loop (decided once per plan from the mesh):
  //   CGF_STRIP2D / CGF_STRIP3D  shuffle-based marching kernels (cg_step2d.cuh / cg_step3d.cuh): 2-D / 3-D meshes with an even m0
  //   CGF_RING                   shared-memory ring k_cg_step (kernels.cuh): everything else (4-D, odd m0); MVTV_STEP=ring forces it
  int cg_family = 0;
  int max_degree = 1;       // highest polynomial degree this plan's kernels implement
  int auto_degree = 1;      // what MVTV_PRECOND_AUTO picks once Jacobi needs more than 24 iterations (measured per family)
  bool fused_update = false;   // one GPU, strip kernels: k_cg_update fused with the first preconditioner pass (r out of place)
  int tune_init3d = 1, tune_defer_rr = 1;   // developer knob MVTV_TUNE: candidates still being measured
  void *r2 = nullptr;          // second residual buffer of the fused update (allocated on first use)
  void *ybuf = nullptr;        // third buffer of the Horner passes, degree >= 3 (allocated on first use; with world > 1 at plan creation)
  int zu_variant = -1;   // ZV_* when the compile-time block tables of k_zu_march match this plan, else -1 (gather kernel)

  // optional per-kernel-class CUDA-event timing on the plan's stream (mvtv_plan_profile)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  struct ProfRec { int cls, e0; };
  std::vector<ProfRec> prof_recs;
  size_t prof_used = 0;
  double prof_ms[MVTV_KC_N] = {0};
  long long prof_cnt[MVTV_KC_N] = {0};
  void prof_begin(int cls) {
    if (!prof_on) return;
    while (prof_ev.size() < prof_used + 2) {
      cudaEvent_t e;
      MVTV_CUDA(cudaEventCreate(&e));
      prof_ev.push_back(e);
    }
    MVTV_CUDA(cudaEventRecord(prof_ev[prof_used], stream));
    prof_recs.push_back({cls, (int)prof_used});
  }
  void prof_end() {
    if (!prof_on) return;
    MVTV_CUDA(cudaEventRecord(prof_ev[prof_used + 1], stream));
    prof_used += 2;
  }
  void prof_flush() {  // call only after a stream synchronize
    if (!prof_on) return;
    for (const ProfRec &r : prof_recs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, prof_ev[r.e0], prof_ev[r.e0 + 1]) == cudaSuccess) {
        prof_ms[r.cls] += ms;
        prof_cnt[r.cls] += 1;
      }
    }
    prof_recs.clear();
    prof_used = 0;
  }

  size_t esz() const { return dtype == MVTV_F64 ? 8 : 4; }

  void use_device() const { MVTV_CUDA(cudaSetDevice(device)); }

  static void *grow(void *&buf, size_t &have, size_t bytes) {
    if (bytes > have) {
      if (buf) MVTV_CUDA(cudaFree(buf));
      buf = nullptr;
      have = 0;
      MVTV_CUDA(cudaMalloc(&buf, bytes));
      have = bytes;
    }
    return buf;
  }

  double *stage(size_t bytes) {
    if (bytes > staging_bytes) {
      if (staging) MVTV_CUDA(cudaFree(staging));
      staging = nullptr;
      staging_bytes = 0;
      MVTV_CUDA(cudaMalloc(&staging, bytes));
      staging_bytes = bytes;
    }
    return staging;
  }

  ~mvtv_plan() {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
    if (comm && world > 1 && !ipc_opened.empty()) {  // nobody frees a buffer a peer still maps
      double *tmp = raw;
      if (tmp && g_nccl.AllReduce) {
        g_nccl.AllReduce(tmp, tmp, 1, ncclFloat64, ncclSum, comm, stream);
        cudaStreamSynchronize(stream);
      }
    }
    if (cb) cudaFree(cb);
    if (d_peer) cudaFree(d_peer);
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    void *bufs[] = {theta, xold, v1, v2, oty, cnt, r, pbuf[0], pbuf[1], q, dinv, zbuf, u[0], u[1], S, zr, raw, partials, counters, vid, staging, in_buf, sort_buf, r2, ybuf};
    for (void *b : bufs)
      if (b) cudaFree(b);
    if (h_scal) cudaFreeHost(h_scal);
    for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }

  // ---- construction ---------------------------------------------------------------------------
  void build_tables(const mvtv_plan_desc &d) {
    p = d.p;
    const int P = p < 2 ? 2 : p;  // a 1-D mesh is stored as (m, 1)
    long long m[MVTV_MAXP];
    for (int a = 0; a < P; ++a) m[a] = a < p ? d.m[a] : 1;
    for (int a = 0; a < p; ++a) MVTV_REQUIRE(m[a] >= 1, "mesh dims must be >= 1");
    N = 1;
    for (int a = 0; a < P; ++a) N *= m[a];
    MVTV_REQUIRE(N < (1ll << 31), "mesh too large for 32-bit in-slab vertex keys");
    dt.P = P;
    long long s = 1;
    for (int a = 0; a < MVTV_MAXP; ++a) {
      dt.m[a] = a < P ? m[a] : 1;
      dt.stride[a] = a < P ? s : 0;
      if (a < P) s *= m[a];
    }
    dt.plane = dt.stride[P - 1];
    // slab partition of the last axis: contiguous, sizes differ by at most one plane
    const long long mz = m[P - 1];
    if (world > 1) {
      if (p < 2) throw Error(MVTV_ERR_UNSUPPORTED, "slab partitioning needs p >= 2");
      if (mz < world) throw Error(MVTV_ERR_INVALID, "last mesh axis shorter than the number of ranks");
    }
    const long long basez = mz / world, extra = mz % world;
    dt.z0 = rank * basez + std::min<long long>(rank, extra);
    dt.nz = (int)(basez + (rank < extra ? 1 : 0));
    dt.has_lo = rank > 0;
    dt.has_hi = rank < world - 1;
    dt.Nloc = dt.plane * dt.nz;
    dt.usz = dt.plane * (dt.nz + 2);

    // difference blocks in create_D order (cpp-code/utils.cpp:245-269)
    const int K = (1 << p) - 1;
    bt.K = K;
    rt.K = K;
    long long off = 0;
    for (int b = 0; b < K; ++b) {
      const int num = (b == 0) ? K : b;  // all-ones mask first (:258), then binaries of 1..K-1 (:259-267)
      int Sm = 0;                        // MSB-first binaries (:73-89): axis a <-> bit p-1-a
      for (int a = 0; a < p; ++a)
        if ((num >> (p - 1 - a)) & 1) Sm |= 1 << a;
      double sc = 1.0;                   // prod_{k not in S} delta_k (:261); all-ones block unscaled
      if (b != 0 && d.deltas)
        for (int a = 0; a < p; ++a)
          if (!((Sm >> a) & 1)) sc *= d.deltas[a];
      int Sp = Sm;
      if (__builtin_popcount(Sm) > 1 && variant == MVTV_VARIANT_REFERENCE) {
        const int lowest = __builtin_ctz(Sm);
        Sp = (Sm & ~(1 << lowest)) | 1;  // first factor along axis 0 (cpp-code/utils.cpp:187)
        if (lowest != 0 && m[0] != m[lowest])
          throw Error(MVTV_ERR_DIM_MISMATCH,
                      "matrix multiplication: incompatible matrix dimensions (reference mixedpartial on a "
                      "non-cubic mesh, cpp-code/utils.cpp:187,216); use MVTV_VARIANT_INTENDED");
      }
      bt.mask[b] = Sp;
      bt.scale[b] = sc;
      bt.nsub[b] = 1 << __builtin_popcount(Sp);
      for (int j = 0; j < bt.nsub[b]; ++j) {
        const int e = pdep_int(j, Sp);
        bt.sub[b][j] = e;
        long long o = 0;
        for (int a = 0; a < P; ++a)
          if ((e >> a) & 1) o += dt.stride[a];
        bt.off[b][j] = o;
      }
      rt.mask[b] = Sp;
      long long rs = 1;
      for (int a = 0; a < MVTV_MAXP; ++a) {
        rt.rstride[b][a] = a < P ? rs : 0;
        if (a < P) rs *= m[a] - ((Sp >> a) & 1);
      }
      rt.rows[b] = rs;
      rt.row_off[b] = off;
      off += rs;
    }
    rt.R = off;

    // compile-time block tables of the marching z/u kernel: use it only if they agree with this plan's table
    {
      const int V = (p == 1) ? ZV_P1 : (variant == MVTV_VARIANT_REFERENCE ? ZV_REFERENCE : ZV_INTENDED);
      bool same = (zu_num_blocks(P, V) == K);
      for (int b = 0; same && b < K; ++b) same = (zu_block_mask(P, V, b) == bt.mask[b]);
      const char *env = getenv("MVTV_ZU_KERNEL");
      zu_variant = (same && !(env && std::string(env) == "gather")) ? V : -1;
    }

    {
      const char *env = getenv("MVTV_STEP");
      const bool ring = env && std::string(env) == "ring";   // MVTV_STEP=ring: the shared-memory k_cg_step everywhere (cross-checks)
      const bool even = (m[0] % 2 == 0) && m[0] >= 2;
      cg_family = (!ring && even && P == 2) ? CGF_STRIP2D : ((!ring && even && P == 3) ? CGF_STRIP3D : CGF_RING);
      const char *efc = getenv("MVTV_FOLD_COMMIT");
      fold_commit = !(efc && std::string(efc) == "0");
    }

    // 3^P-point stencil of D^T D = sum_b c_b^2 kron_{a in S'_b} L_a  (SURVEY A.6), clamped indices
    st.npts = 1;
    for (int a = 0; a < P; ++a) st.npts *= 3;
    const double t3[3] = {-1.0, 2.0, -1.0};
    for (int o = 0; o < st.npts; ++o) {
      double c = 0.0;
      for (int b = 0; b < K; ++b) {
        double w = bt.scale[b] * bt.scale[b];
        int rem = o;
        for (int a = 0; a < P; ++a) {
          const int dgt = rem % 3;
          rem /= 3;
          if ((bt.mask[b] >> a) & 1) w *= t3[dgt];
          else w *= (dgt == 1) ? 1.0 : 0.0;
        }
        c += w;
      }
      st.coef[o] = c;
    }
    for (int cls = 0; cls < (1 << P); ++cls) {
      double dsum = 0.0;
      for (int b = 0; b < K; ++b) {
        double w = bt.scale[b] * bt.scale[b];
        for (int a = 0; a < P; ++a)
          if ((bt.mask[b] >> a) & 1) {
            const bool bnd = (cls >> a) & 1;
            w *= bnd ? (m[a] >= 2 ? 1.0 : 0.0) : 2.0;
          }
        dsum += w;
      }
      st.diagK[cls] = dsum;
    }
    // Gershgorin bound on spec(D^-1 M), M = diag(c) + s K: the row of K at a vertex of boundary class cls is the
    // stencil with the offsets that leave the mesh folded back onto the clamped neighbour; (c + s*rowabs)/(c + s*diag)
    // is largest at c = 0, so the bound is max_cls rowabs/diag whatever c and s are.
    cheb_bmax = 1.0;
    for (int cls = 0; cls < (1 << P); ++cls) {
      bool possible = true;
      for (int a = 0; a < P; ++a)
        if (!((cls >> a) & 1) && m[a] < 3) possible = false;   // an interior vertex needs m >= 3 on that axis
      if (!possible) continue;
      std::vector<double> row(st.npts, 0.0);
      for (int o = 0; o < st.npts; ++o) {
        int rem = o, tgt = 0, mul = 1;
        for (int a = 0; a < P; ++a) {
          int dgt = rem % 3;
          rem /= 3;
          if ((cls >> a) & 1) {
            if (m[a] < 2) dgt = 1;                 // extent-1 axis: every neighbour is the vertex itself
            else if (dgt == 0) dgt = 1;            // low boundary: the -1 neighbour folds onto the vertex
          }
          tgt += dgt * mul;
          mul *= 3;
        }
        row[tgt] += st.coef[o];
      }
      int centre = 0, mul = 1;
      for (int a = 0; a < P; ++a) { centre += mul; mul *= 3; }
      double rowabs = 0.0;
      for (double v : row) rowabs += fabs(v);
      if (row[centre] > 0.0) cheb_bmax = std::max(cheb_bmax, rowabs / row[centre]);
    }
  }

  void allocate() {
    use_device();
    MVTV_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    MVTV_CUDA(cudaEventCreate(&ev0));
    MVTV_CUDA(cudaEventCreate(&ev1));
    const size_t vb = (size_t)dt.usz * esz();
    void **vecs[] = {&theta, &xold, &v1, &v2, &oty, &cnt, &r, &pbuf[0], &pbuf[1], &q, &dinv, &zbuf};
    for (void **v : vecs) {
      MVTV_CUDA(cudaMalloc(v, vb));
      MVTV_CUDA(cudaMemsetAsync(*v, 0, vb, stream));
    }
    for (int k = 0; k < 2; ++k) {
      MVTV_CUDA(cudaMalloc(&u[k], vb * bt.K));
      MVTV_CUDA(cudaMemsetAsync(u[k], 0, vb * bt.K, stream));
    }
    block = dim3(256, 1, 1);
    grid = dim3((unsigned)((dt.plane + 255) / 256), (unsigned)(dt.nz + dt.has_lo), 1);
    MVTV_REQUIRE(grid.y <= 65535, "last mesh axis too long for grid.y");
    nblocks = grid.x * grid.y;
    MVTV_CUDA(cudaMalloc(&partials, sizeof(double) * (size_t)std::max<unsigned>(nblocks, 1u << 16) * ZR_N));
    MVTV_CUDA(cudaMalloc(&counters, sizeof(unsigned) * 8));
    MVTV_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned) * 8, stream));
    MVTV_CUDA(cudaMalloc(&S, sizeof(double) * CS_N));
    MVTV_CUDA(cudaMalloc(&zr, sizeof(double) * 8));
    MVTV_CUDA(cudaMalloc(&raw, sizeof(double) * 8));
    MVTV_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * CS_N, stream));
    MVTV_CUDA(cudaMemsetAsync(zr, 0, sizeof(double) * 8, stream));
    MVTV_CUDA(cudaMemsetAsync(raw, 0, sizeof(double) * 8, stream));
    MVTV_CUDA(cudaMallocHost(&h_scal, sizeof(double) * 16));
    MVTV_CUDA(cudaStreamSynchronize(stream));
  }

  // Which preconditioner degrees and which update kernel this plan runs; called once the collectives are known.
  // Horner passes of degree >= 2 and the fused update exist in the strip kernels: on one GPU, and on several GPUs for 3-D
  // meshes over the peer-memory path with folded commits (their halo rows need q / w of the neighbour rank, which only
  // k_cg_step3d exchanges).
  void finalize_selection() {
    const bool multi3 = world > 1 && cg_family == CGF_STRIP3D && d_peer && fold_commit;
    const bool full = (cg_family != CGF_RING) && (world == 1 || multi3);
    max_degree = full ? 4 : 1;
    auto_degree = full ? 3 : 1;
    fused_update = full;
    // MVTV_TUNE="key=value,...": developer knob for A/B measurements of candidates that are still compiled in
    // (fused=0|1, degree=1..4 for MVTV_PRECOND_AUTO, init3d=0|1: gather / marching CG initialisation, defer_rr=0|1, kappa=<interval ratio of the polynomial>)
    if (const char *tune = getenv("MVTV_TUNE")) {
      std::string t(tune);
      size_t pos = 0;
      while (pos < t.size()) {
        size_t end = t.find(',', pos);
        if (end == std::string::npos) end = t.size();
        const std::string kv = t.substr(pos, end - pos);
        const size_t eq = kv.find('=');
        if (eq != std::string::npos) {
          const std::string k = kv.substr(0, eq);
          const int v = atoi(kv.c_str() + eq + 1);
          if (k == "fused") fused_update = fused_update && v != 0;
          else if (k == "degree") auto_degree = std::max(1, std::min(v, max_degree));
          else if (k == "init3d") tune_init3d = v;
          else if (k == "defer_rr") tune_defer_rr = v;
          else if (k == "kappa") cheb_kappa = std::max(2, v);
        }
        pos = end + 1;
      }
    }
  }

  dim3 grid_owned() const { return dim3(grid.x, (unsigned)dt.nz, 1); }
  static int grid1d(long long n) {
    long long g = (n + 255) / 256;
    return (int)std::max<long long>(1, std::min<long long>(g, 148 * 16));
  }

  // ---- multi-GPU plumbing --------------------------------------------------------------------
  template <typename T>
  void exchange_ghosts(T *x) {
    if (world == 1) return;
    const size_t pl = (size_t)dt.plane;
    MVTV_NCCL(g_nccl.GroupStart());
    if (dt.has_hi) {
      MVTV_NCCL(g_nccl.Send(x + (size_t)dt.nz * pl, pl, NcclType<T>::v, rank + 1, comm, stream));
      MVTV_NCCL(g_nccl.Recv(x + (size_t)(dt.nz + 1) * pl, pl, NcclType<T>::v, rank + 1, comm, stream));
    }
    if (dt.has_lo) {
      MVTV_NCCL(g_nccl.Send(x + pl, pl, NcclType<T>::v, rank - 1, comm, stream));
      MVTV_NCCL(g_nccl.Recv(x, pl, NcclType<T>::v, rank - 1, comm, stream));
    }
    MVTV_NCCL(g_nccl.GroupEnd());
  }
  void allreduce(double *buf, int n, ncclRedOp_t op) {
    if (world == 1) return;
    MVTV_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, op, comm, stream));
  }

  // Share the comm buffer and r with the other ranks through CUDA IPC and build the device peer table.
  void setup_peer() {
    const char *env = getenv("MVTV_COMM");
    if (world < 2 || world > MVTV_PEER_MAXW || (env && std::string(env) == "nccl")) return;
    const size_t n_slots = (size_t)MVTV_PEER_NSLOT * MVTV_PEER_MAXW * MVTV_PEER_NVAL;   // doubles
    const size_t n_flags = (size_t)MVTV_PEER_NSLOT * MVTV_PEER_MAXW;                    // u64
    const size_t cb_bytes = 8 * (n_slots + n_flags + 4 + 1);
    MVTV_CUDA(cudaMalloc(&cb, cb_bytes));
    MVTV_CUDA(cudaMemset(cb, 0, cb_bytes));
    // the 3-D strip kernels also exchange q and the buffers of the Horner passes (z, the two direction buffers, y)
    const size_t vbytes = (size_t)dt.usz * esz();
    if (cg_family == CGF_STRIP3D && !ybuf) {
      MVTV_CUDA(cudaMalloc(&ybuf, vbytes));
      MVTV_CUDA(cudaMemset(ybuf, 0, vbytes));
    }
    struct Handles { cudaIpcMemHandle_t cb, r, z, q, p0, p1, y; };
    static_assert(sizeof(Handles) == 448, "seven 64-byte IPC handles");
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    bool ok = cudaIpcGetMemHandle(&mine.cb, cb) == cudaSuccess && cudaIpcGetMemHandle(&mine.r, r) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.z, zbuf) == cudaSuccess && cudaIpcGetMemHandle(&mine.q, q) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.p0, pbuf[0]) == cudaSuccess && cudaIpcGetMemHandle(&mine.p1, pbuf[1]) == cudaSuccess &&
              (!ybuf || cudaIpcGetMemHandle(&mine.y, ybuf) == cudaSuccess);
    // all ranks must take the same decision: all-reduce the ok flag with the handles' exchange
    unsigned char *d_h = nullptr;
    MVTV_CUDA(cudaMalloc(&d_h, sizeof(Handles) * world + 8));
    std::vector<Handles> all(world);
    MVTV_CUDA(cudaMemcpy(d_h + sizeof(Handles) * rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice));
    MVTV_NCCL(g_nccl.AllGather(d_h + sizeof(Handles) * rank, d_h, sizeof(Handles), ncclChar, comm, stream));
    MVTV_CUDA(cudaStreamSynchronize(stream));
    MVTV_CUDA(cudaMemcpy(all.data(), d_h, sizeof(Handles) * world, cudaMemcpyDeviceToHost));
    std::vector<unsigned char *> pcb(world, nullptr);
    // neighbours' buffers, in the order r, z, q, p0, p1, y
    unsigned char *nb_prev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, *nb_next[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int j = 0; ok && j < world; ++j) {
      if (j == rank) { pcb[j] = cb; continue; }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[j].cb, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      ipc_opened.push_back(p);
      pcb[j] = (unsigned char *)p;
      if (j == rank - 1 || j == rank + 1) {
        const cudaIpcMemHandle_t hs[6] = {all[j].r, all[j].z, all[j].q, all[j].p0, all[j].p1, all[j].y};
        for (int k = 0; ok && k < (ybuf ? 6 : 5); ++k) {
          void *pb = nullptr;
          if (cudaIpcOpenMemHandle(&pb, hs[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
          ipc_opened.push_back(pb);
          (j == rank - 1 ? nb_prev : nb_next)[k] = (unsigned char *)pb;
        }
      }
    }
    cudaGetLastError();
    double flag = ok ? 0.0 : 1.0;   // anybody failed -> everybody uses NCCL
    MVTV_CUDA(cudaMemcpy(d_h, &flag, 8, cudaMemcpyHostToDevice));
    MVTV_NCCL(g_nccl.AllReduce(d_h, d_h, 1, ncclFloat64, ncclSum, comm, stream));
    MVTV_CUDA(cudaStreamSynchronize(stream));
    MVTV_CUDA(cudaMemcpy(&flag, d_h, 8, cudaMemcpyDeviceToHost));
    MVTV_CUDA(cudaFree(d_h));
    if (flag != 0.0) {
      for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
      ipc_opened.clear();
      return;
    }
    PeerTab pt{};
    pt.rank = rank;
    pt.world = world;
    pt.has_lo = dt.has_lo;
    pt.has_hi = dt.has_hi;
    auto slots_of = [&](unsigned char *b) { return (double *)b; };
    auto flags_of = [&](unsigned char *b) { return (unsigned long long *)(b + 8 * n_slots); };
    auto hflags_of = [&](unsigned char *b) { return (unsigned long long *)(b + 8 * (n_slots + n_flags)); };
    for (int j = 0; j < world; ++j) { pt.slots[j] = slots_of(pcb[j]); pt.flags[j] = flags_of(pcb[j]); }
    // hflags[0] = "my ghost plane BELOW was filled by the previous rank", hflags[1] = "... ABOVE by the next rank"
    pt.hflag_from_prev = hflags_of(cb) + 0;
    pt.hflag_from_next = hflags_of(cb) + 1;
    pt.hflag_at_prev = dt.has_lo ? hflags_of(pcb[rank - 1]) + 1 : nullptr;
    pt.hflag_at_next = dt.has_hi ? hflags_of(pcb[rank + 1]) + 0 : nullptr;
    // the slab of rank j has its own nz; its ghost-above plane starts at (nz_j + 1) * plane
    const long long mz = dt.m[dt.P - 1], basez = mz / world, extra = mz % world;
    auto nz_of = [&](int j) { return basez + (j < extra ? 1 : 0); };
    // rank j-1's ghost plane ABOVE its slab starts at (nz_{j-1} + 1) * plane; rank j+1's ghost plane BELOW is its plane 0
    auto at_prev = [&](int k) { return (dt.has_lo && nb_prev[k]) ? (void *)(nb_prev[k] + esz() * (size_t)((nz_of(rank - 1) + 1) * dt.plane)) : nullptr; };
    auto at_next = [&](int k) { return (dt.has_hi && nb_next[k]) ? (void *)nb_next[k] : nullptr; };
    pt.rghost_at_prev = at_prev(0);
    pt.rghost_at_next = at_next(0);
    pt.zghost_at_prev = at_prev(1);
    pt.zghost_at_next = at_next(1);
    pt.qghost_at_prev = at_prev(2);
    pt.qghost_at_next = at_next(2);
    const int wk[4] = {1, 3, 4, 5};   // PW_Z, PW_P0, PW_P1, PW_Y
    for (int w = 0; w < 4; ++w) { pt.wghost_at_prev[w] = at_prev(wk[w]); pt.wghost_at_next[w] = at_next(wk[w]); }
    pt.zflag_from_prev = hflags_of(cb) + 2;
    pt.zflag_from_next = hflags_of(cb) + 3;
    pt.zflag_at_prev = dt.has_lo ? hflags_of(pcb[rank - 1]) + 3 : nullptr;
    pt.zflag_at_next = dt.has_hi ? hflags_of(pcb[rank + 1]) + 2 : nullptr;
    pt.error = (int *)(cb + 8 * (n_slots + n_flags + 4));
    MVTV_CUDA(cudaMalloc(&d_peer, sizeof(PeerTab)));
    MVTV_CUDA(cudaMemcpy(d_peer, &pt, sizeof(PeerTab), cudaMemcpyHostToDevice));
  }

  // ---- points --------------------------------------------------------------------------------
  template <typename T>
  void set_points_t(long long npts, const double *data_dev, long long ld_point, long long ld_axis,
                    const double *y_dev, const double *axes_dev) {
    use_device();
    // world > 1: a rank whose slab holds no points (clustered data, strong scaling) still takes part in every collective
    MVTV_REQUIRE(npts >= (world > 1 ? 0 : 1) && npts < (1ll << 31), "n must be in [1, 2^31)");
    if (npts > vid_cap) {
      if (vid) MVTV_CUDA(cudaFree(vid));
      vid = nullptr;
      vid_cap = 0;
      MVTV_CUDA(cudaMalloc(&vid, sizeof(long long) * (size_t)npts));
      vid_cap = npts;
    }
    have_points = false;   // a failure below must not leave a half-built operator behind
    const size_t vb = (size_t)dt.usz * esz();
    MVTV_CUDA(cudaMemsetAsync(oty, 0, vb, stream));
    MVTV_CUDA(cudaMemsetAsync(cnt, 0, vb, stream));
    double foreign = 0.0;
    if (npts > 0) {
      const size_t tb = sort_temp_bytes(npts);
      const size_t kb = (sizeof(unsigned) * (size_t)npts * 4 + 255) & ~(size_t)255;
      unsigned *key_in = (unsigned *)grow(sort_buf, sort_bytes, kb + std::max<size_t>(tb, 16));
      unsigned *key_out = key_in + npts, *val_in = key_out + npts, *val_out = val_in + npts;
      void *temp = (unsigned char *)key_in + kb;
      launch_bin(p, dt, npts, data_dev, ld_point, ld_axis, axes_dev, vid, key_in, val_in, stream);
      launch_sort(temp, tb, npts, key_in, key_out, val_in, val_out, stream);
      launch_segment_reduce<T>(npts, key_out, val_out, y_dev, dt.plane, (T *)oty, (T *)cnt, stream);
      launches += 3;
      if (world > 1) {
        // every rank must hold exactly the points of its own slab (partition.exchange_points); the verdict is
        // all-reduced below so that no rank throws while its peers wait in a collective
        unsigned last_key = 0;
        MVTV_CUDA(cudaMemcpyAsync(&last_key, key_out + (npts - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        MVTV_CUDA(cudaStreamSynchronize(stream));
        if (last_key == 0xFFFFFFFFu) foreign = 1.0;
      }
    }
    // mean(y) (cpp-code/solvers.cpp:95,103): deterministic two-level sum on the device; world > 1: the point count and the
    // foreign-point flag ride on the same all-reduce
    const double flagged = sum_y(y_dev, npts, foreign);
    if (flagged != 0.0)
      throw Error(MVTV_ERR_INVALID, "world > 1: a point's nearest vertex lies outside its rank's slab (on at least one rank)");
    exchange_ghosts<T>((T *)cnt);   // diag(c) on the ghost planes feeds the redundant ghost-plane work
    MVTV_CUDA(cudaStreamSynchronize(stream));
    dinv_rho = NAN;
    n = npts;
    have_points = true;
    have_u_state = have_theta_state = false;
    jac_estimate = 100.0;   // new operators: MVTV_PRECOND_AUTO starts over (results then depend on the data only, not on
    last_cg_iters = 8;      // what the plan solved before)
    last_cg_prec = 0;
  }
  double sum_y(const double *y_dev, long long npts, double flag);

  // ---- the ADMM loop ---------------------------------------------------------------------------
  template <typename T, int P>
  int cg_solve(double rho, double usc, double rhoM, double rtol, int maxit, int prec, long long &inner, int &status);
  template <typename T>
  void launch_zu(double kappa, double usc, int mode, int init, bool with_prev);
  template <typename T, typename Cfg, int V>
  void launch_zu_march(const ZuArgs<T> &a, const RedBuf &rb);
  template <typename T>
  int solve_t(const mvtv_solve_params &prm, const double *theta_init, double *u_inout, double *theta_out,
              double *fitted_out, mvtv_solve_result &res);
  template <typename T, int P>
  int lambda_max_t(int mode, double *lam_out, int *iters_out);
  template <typename T>
  void read_scalars(const double *dev, int count) {
    MVTV_CUDA(cudaMemcpyAsync(h_scal, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
    MVTV_CUDA(cudaStreamSynchronize(stream));
    prof_flush();
  }
};

// sum of y with the deterministic grid reduction
__global__ void __launch_bounds__(256) k_sum(const double *__restrict__ y, long long n, RedBuf rb, double *out) {
  double red[1] = {0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    red[0] += y[i];
  grid_reduce<1, 1>(red, rb, [out](const double (&res)[1]) { out[0] = res[0]; });
}
// max |x - c| over the owned slab (the first CPP/PY loop test against thetaold = mean(y)-0.1, cpp :103-104,113)
template <typename T>
__global__ void __launch_bounds__(256) k_maxabs_const(const T *__restrict__ x_ghosted, long long plane, long long nloc,
                                                      double c, RedBuf rb, double *out) {
  double red[1] = {0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (long long)gridDim.x * blockDim.x)
    red[0] = fmax(red[0], fabs((double)x_ghosted[plane + i] - c));
  grid_reduce<1, 0>(red, rb, [out](const double (&res)[1]) { out[0] = res[0]; });
}

// sum_i (theta[vertex(i)] - target_i)^2 over this rank's points: the numerator of mse() (cpp-code/solvers.cpp:160-163)
template <typename T>
__global__ void __launch_bounds__(256) k_sqerr(const long long *__restrict__ vid, const T *__restrict__ theta_ghosted,
                                               const double *__restrict__ target, long long n, long long plane,
                                               long long z0, RedBuf rb, double *out) {
  double red[1] = {0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = (double)theta_ghosted[plane + (vid[i] - z0 * plane)] - target[i];
    red[0] += d * d;
  }
  grid_reduce<1, 1>(red, rb, [out](const double (&res)[1]) { out[0] = res[0]; });
}


// ---- small generic kernels for the one-off lambda_max estimate (cpp-code/utils.cpp:354-404) -------------------
template <typename T>
__global__ void __launch_bounds__(256) k_axpby(T *out, double a, const T *x, double b, const T *y, long long plane,
                                               long long nloc) {  // out = a*x + b*y on the owned slab
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (long long)gridDim.x * blockDim.x)
    out[plane + i] = (T)(a * (double)x[plane + i] + b * (double)y[plane + i]);
}
template <typename T>
__global__ void __launch_bounds__(256) k_dot(const T *x, const T *y, long long plane, long long nloc, RedBuf rb,
                                             double *out) {
  double red[1] = {0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (long long)gridDim.x * blockDim.x)
    red[0] += (double)x[plane + i] * (double)y[plane + i];
  grid_reduce<1, 1>(red, rb, [out](const double (&res)[1]) { out[0] = res[0]; });
}
// max over the rows of D of |(D x)_row|   (lam_max_pinv: max(abs(a*b)), cpp-code/utils.cpp:399-404)
template <typename T>
__global__ void __launch_bounds__(256) k_maxabs_D(const __grid_constant__ DimTab dt, const __grid_constant__ BlockTab bt,
                                                  const T *x, RedBuf rb, double *out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zl = blockIdx.y;
  double red[1] = {0.0};
  if (q < dt.plane) {
    int lo_ok, hi_ok;
    boundary_masks(dt, q, zl, lo_ok, hi_ok);
    const long long base = (long long)(zl + 1) * dt.plane + q;
    for (int b = 0; b < bt.K; ++b) {
      const int S = bt.mask[b];
      if ((S & ~hi_ok) != 0) continue;
      double d = 0.0;
      for (int f = 0; f < bt.nsub[b]; ++f) {
        const double t = (double)x[base + bt.off[b][f]];
        d += (__popc(bt.sub[b][f]) & 1) ? -t : t;
      }
      red[0] = fmax(red[0], fabs(bt.scale[b] * d));
    }
  }
  grid_reduce<1, 0>(red, rb, [out](const double (&res)[1]) { out[0] = res[0]; });
}

// returns the world-wide sum of `flag` (set_points' error word); sets mean_y
double mvtv_plan::sum_y(const double *y_dev, long long npts, double flag) {
  const int g = std::min<int>(grid1d(npts), (int)nblocks);
  RedBuf rb{partials, counters + 4};
  k_sum<<<g, 256, 0, stream>>>(y_dev, npts, rb, zr);
  MVTV_CUDA(cudaGetLastError());
  launches += 1;
  double extra[2] = {(double)npts, flag};
  if (world > 1) {  // global mean over the disjoint per-rank point sets
    MVTV_CUDA(cudaMemcpyAsync(zr + 1, extra, sizeof(double) * 2, cudaMemcpyHostToDevice, stream));
    allreduce(zr, 3, ncclSum);
  }
  MVTV_CUDA(cudaMemcpyAsync(h_scal, zr, sizeof(double) * 3, cudaMemcpyDeviceToHost, stream));
  MVTV_CUDA(cudaStreamSynchronize(stream));
  const double cnt_total = world > 1 ? h_scal[1] : (double)npts;
  const double flagged = world > 1 ? h_scal[2] : flag;
  if (flagged == 0.0) {
    MVTV_REQUIRE(cnt_total >= 1.0, "no points on any rank");
    mean_y = h_scal[0] / cnt_total;
  }
  return flagged;
}

template <typename T, typename Cfg, int V>
void mvtv_plan::launch_zu_march(const ZuArgs<T> &a, const RedBuf &rb) {
  constexpr int Q = Cfg::Q;
  const size_t smem = sizeof(T) * (size_t)Cfg::SMEM_ELEMS;
  // function attributes are per device: cache them per ordinal (one process may own plans on several GPUs)
  static bool attr_set_dev[64] = {false};
  static int occ_dev[64];
  const int di = device & 63;
  if (!attr_set_dev[di]) {
    MVTV_CUDA(cudaFuncSetAttribute(k_zu_march<T, Cfg, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MVTV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dev[di], k_zu_march<T, Cfg, V>, Cfg::NT, smem));
    if (occ_dev[di] < 1) occ_dev[di] = 1;
    attr_set_dev[di] = true;
  }
  const int occ = occ_dev[di];
  const int m0 = (int)dt.m[0], m1 = Q >= 2 ? (int)dt.m[1] : 1, m2 = Q >= 3 ? (int)dt.m[2] : 1;
  const long long tiles = (long long)((m0 + Cfg::OX - 1) / Cfg::OX) * ((m1 + Cfg::OY - 1) / Cfg::OY) *
                          ((m2 + Cfg::OW - 1) / Cfg::OW);
  int nsm = 148;
  MVTV_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  const long long slots = (long long)nsm * occ;
  int nchunk = 1;
  double best = -1.0;
  const int maxchunk = std::max(1, std::min(dt.nz / 16, 256));
  for (int nc = 1; nc <= maxchunk; ++nc) {
    const int zc = (dt.nz + nc - 1) / nc;
    const int ncr = (dt.nz + zc - 1) / zc;
    const long long total = tiles * ncr;
    if (total > (1ll << 16)) break;
    const long long waves = (total + slots - 1) / slots;
    const double eff = (double)total / (double)(waves * slots) * ((double)zc / (zc + 1.0));
    if (eff > best + 1e-9) { best = eff; nchunk = ncr; }
  }
  const int zchunk = (dt.nz + nchunk - 1) / nchunk;
  nchunk = (dt.nz + zchunk - 1) / zchunk;
  k_zu_march<T, Cfg, V><<<dim3((unsigned)tiles, (unsigned)nchunk, 1), Cfg::NT, smem, stream>>>(dt, bt, a, rb, zchunk);
}

template <typename T>
void mvtv_plan::launch_zu(double kappa, double usc, int mode, int init, bool with_prev) {
  ZuArgs<T> a;
  a.theta = (const T *)theta;
  a.theta_prev = with_prev ? (const T *)xold : nullptr;
  a.u_old = (const T *)u[ucur];
  a.u_new = (T *)u[ucur ^ 1];
  a.v1 = (T *)v1;
  a.v2 = (T *)v2;
  a.kappa = kappa;
  a.uscale = usc;
  a.mode = mode;
  a.init = init;
  a.red_out = zr;
  RedBuf rb{partials, counters + 0};
  prof_begin(init ? MVTV_KC_ZU_INIT : MVTV_KC_ZU);
  bool done = false;
  if (zu_variant >= 0) {
    switch (dt.P * 4 + zu_variant) {
      case 2 * 4 + ZV_REFERENCE: launch_zu_march<T, ZuCfg<2, 256, 1, 1>, ZV_REFERENCE>(a, rb); done = true; break;
      case 2 * 4 + ZV_INTENDED: launch_zu_march<T, ZuCfg<2, 256, 1, 1>, ZV_INTENDED>(a, rb); done = true; break;
      case 2 * 4 + ZV_P1: launch_zu_march<T, ZuCfg<2, 256, 1, 1>, ZV_P1>(a, rb); done = true; break;
      // tiles from the measured sweep (profiles/r2_call1_zu_probe.log): 3-D 16x16 (6.11 against 6.39 ms per launch on 512^3),
      // 4-D 8x8x4 with 256 threads and no register spills (17.9 against 21.2 ms on 96^4)
      case 3 * 4 + ZV_REFERENCE: launch_zu_march<T, ZuCfg<3, 16, 16, 1>, ZV_REFERENCE>(a, rb); done = true; break;
      case 3 * 4 + ZV_INTENDED: launch_zu_march<T, ZuCfg<3, 16, 16, 1>, ZV_INTENDED>(a, rb); done = true; break;
      case 4 * 4 + ZV_REFERENCE: launch_zu_march<T, ZuCfg<4, 8, 8, 4>, ZV_REFERENCE>(a, rb); done = true; break;
      case 4 * 4 + ZV_INTENDED: launch_zu_march<T, ZuCfg<4, 8, 8, 4>, ZV_INTENDED>(a, rb); done = true; break;
      default: break;
    }
  }
  if (!done) k_zu<T><<<grid, block, 0, stream>>>(dt, bt, a, rb);
  prof_end();
  MVTV_CUDA(cudaGetLastError());
  launches += 1;
  if (world > 1) {
    allreduce(zr, ZR_NSUM, ncclSum);
    allreduce(zr + ZR_NSUM, ZR_N - ZR_NSUM, ncclMax);
  }
}

// tile shapes of k_cg_step per mesh rank (in-plane tile x planes marched by one CTA)
template <int P> struct StepShape;
//                                                  Q  TXT XO  TY RY TW DEPTH
template <> struct StepShape<2> { using Cfg = StepCfg<1, 256, 2,  1, 1, 1, 3>; };   // 512-wide rows, 256 threads, 41 KB
template <> struct StepShape<3> { using Cfg = StepCfg<2,  32, 1, 16, 4, 1, 4>; };   // 32x16 tile, 128 threads, 64 KB
template <> struct StepShape<4> { using Cfg = StepCfg<3,  32, 1,  8, 2, 4, 3>; };   // 32x8x4 tile, 512 threads, 163 KB

// Tile variants of the strip kernels, chosen from the measured sweeps (profiles/r1_final2_step2d_probe*.log, r2_call1_*.log):
//   2-D (4096^2): direction + SpMV strips of 128 (4 vertices per lane) 111 us; first preconditioner pass 64 registers, 32 warps
//       per SM, diag(c) derived from dinv 84 us; Horner passes 3 CTAs of 8 warps per SM; fused update 8 warps
//   3-D (512^3): direction + SpMV 4 warps x 4 rows 1.02 ms; preconditioner passes 4 warps x 3 rows 0.78 ms
using S2Step = Step2dCfg<4, 1, 4, 0>;
using S2Prec = Step2dCfg<8, 1, 2, 4, true>;
using S2Fuse = Fused2dCfg<8, 0>;
constexpr int S2_HORNER_WARPS = 8, S2_HORNER_MINB = 3, S2_INIT_WARPS = 4;
// tiles of the 3-D strip kernels.  TMA staging (Step3dCfg::TMAD = 3: cp.async.bulk row copies into a per-warp ring) was measured
// on every mode (profiles/r2d_probe3_tma.log, 512^3): it loses 5 / 13 / 25 % on the direction, fused-update and Horner kernels --
// their rows are consumed once, the extra shared-memory hop and the 512-byte copy granularity cost more than the deeper prefetch
// gains -- and wins on the CG initialisation (2.0 against 3.0 ms), whose five own-row side streams compete with the staged rows
// for registers; so only that mode uses it.
struct S3Tiles {
  using Step = Step3dCfg<4, 4>;
  using Prec = Step3dCfg<4, 3>;                    // also the Horner passes
  using Fuse = Step3dCfg<4, 2>;                    // 4 x 3 rows spills (255 registers), 8 x 2 measures the same
  using Init = Step3dCfg<4, 3, 0, true, 3>;        // TMA staging of theta's rows
};

template <typename T, int P>
int mvtv_plan::cg_solve(double rho, double usc, double rhoM, double rtol, int maxit, int deg, long long &inner, int &status) {
  using Cfg = typename StepShape<P>::Cfg;
  constexpr int Q = P - 1;
  const int fam = cg_family;
  // the family is a run-time property of the plan, the kernels are templates on P: guard the instantiations
  constexpr bool HAS2D = (P == 2), HAS3D = (P == 3);
  deg = std::max(0, std::min(deg, max_degree));
  if (!(dinv_rho == rhoM)) {
    k_make_dinv<T, P><<<dim3(grid.x, (unsigned)dt.nz + 2, 1), block, 0, stream>>>(dt, st, (const T *)cnt, rhoM, (T *)dinv);
    MVTV_CUDA(cudaGetLastError());
    dinv_rho = rhoM;
    launches += 1;
  }
  const bool fused = fused_update && deg >= 1;   // update + first preconditioner pass in one kernel, r out of place
  if (fused && !r2) {
    MVTV_CUDA(cudaMalloc(&r2, (size_t)dt.usz * esz()));
    MVTV_CUDA(cudaMemsetAsync(r2, 0, (size_t)dt.usz * esz(), stream));
  }
  if (deg >= 3 && !ybuf) {   // world > 1: allocated (and shared with the neighbours) at plan creation
    MVTV_CUDA(cudaMalloc(&ybuf, (size_t)dt.usz * esz()));
    MVTV_CUDA(cudaMemsetAsync(ybuf, 0, (size_t)dt.usz * esz(), stream));
  }
  CgArgs<T> a;
  a.x = (T *)theta;
  a.xold = (T *)xold;
  a.r = (T *)r;
  a.r2 = fused ? (T *)r2 : nullptr;
  a.q = (T *)q;
  a.pbuf[0] = (T *)pbuf[0];
  a.pbuf[1] = (T *)pbuf[1];
  a.c = (const T *)cnt;
  a.dinv = (const T *)dinv;
  a.oty = (const T *)oty;
  a.v1 = (const T *)v1;
  a.v2 = (const T *)v2;
  a.S = S;
  a.raw = (world > 1 && !d_peer) ? raw : nullptr;
  a.peer = d_peer;
  const bool fold = d_peer && fold_commit;   // commits folded into the reducing kernels (no k_cg_peer_commit_* launches)
  a.fold = fold ? 1 : 0;
  a.seq_red = 0;
  a.seq_halo = 0;
  a.seq_zhalo = 0;
  a.z = (T *)zbuf;
  a.y = (T *)ybuf;
  a.prec = deg;
  // buffer a preconditioner pass writes (0 = z, 1 = idle direction buffer, 2 = y): no buffer is rewritten before the pass
  // after next, the last pass writes z
  auto wsel = [deg](int j) {
    static const int rot[5][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {1, 0, 0, 0}, {1, 2, 0, 0}, {0, 1, 2, 0}};
    return rot[deg][j - 1];
  };
  unsigned long long z_event = 0;   // z-flag event of the latest exported planes (several GPUs)
  unsigned long long rr_pending = 0;   // reduction event of a fused update whose r.r the last Horner pass still has to collect
  // coefficients c_0 .. c_d of the degree-d polynomial P in D^-1 M whose residual 1 - t P(t) is the Chebyshev polynomial
  // T_{d+1} on [bmax/kappa, bmax]; Horner form: w_1 = c_d A z0 + c_{d-1} z0, w_k = A w_{k-1} + c_{d-k} z0, z = w_d
  double hc[8] = {0};
  if (deg >= 1) cheb_poly_coeffs(deg, cheb_bmax, cheb_kappa, hc);
  a.pc0 = deg >= 1 ? hc[deg - 1] : 0.0;
  a.pc1 = deg >= 1 ? hc[deg] : 0.0;
  a.rho = rho;
  a.uscale = usc;
  a.rhoM = rhoM;
  a.rtol2 = rtol * rtol;
  int nsm = 148;
  MVTV_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  const int m0 = (int)dt.m[0], m1 = Q >= 2 ? (int)dt.m[1] : 1, m2 = Q >= 3 ? (int)dt.m[2] : 1;
  // launch shape of a marching kernel: in-plane tiles x chunks of the last axis.  Chunks are sized per kernel (their register
  // counts, hence resident CTAs per SM, differ) to fill whole waves of (SMs x resident CTAs) while keeping chunks >= 16 planes
  // so the two extra planes a chunk stages stay cheap.
  auto chunking = [&](unsigned tiles, int occ, int &zchunk_out) {
    const long long slots = (long long)nsm * std::max(occ, 1);
    int nchunk = 1;
    double best = -1.0;
    const int maxchunk = std::max(1, std::min(dt.nz / 16, 4096));
    for (int nc = 1; nc <= maxchunk; ++nc) {
      const int zc = (dt.nz + nc - 1) / nc;
      const int ncr = (dt.nz + zc - 1) / zc;
      const long long total = (long long)tiles * ncr;
      if (total > (1ll << 16)) break;
      const long long waves = (total + slots - 1) / slots;
      const double eff = (double)total / (double)(waves * slots) * ((double)zc / (zc + 2.0));
      if (eff > best + 1e-9) { best = eff; nchunk = ncr; }
    }
    zchunk_out = (dt.nz + nchunk - 1) / nchunk;
    nchunk = (dt.nz + zchunk_out - 1) / zchunk_out;
    return dim3(tiles, (unsigned)nchunk, 1);
  };
  auto with3 = [](auto &&fn) { fn(S3Tiles{}); };   // run fn with the 3-D tile set as a type
  // the marching CG initialisation stages rows with 16-byte bulk copies: fp32 meshes need m0 % 4 == 0
  const bool init3d = (fam == CGF_STRIP3D) && tune_init3d && ((dt.m[0] * (long long)esz()) % 16 == 0);
  // occupancy queries are per device and per kernel: cache them per (device, family, slot)
  struct Shapes { bool set = false; int occ[8] = {1, 1, 1, 1, 1, 1, 1, 1}; };
  static Shapes shapes_dev[64][3];
  Shapes &sh = shapes_dev[device & 63][fam];
  enum { K_STEP_J = 0, K_STEP_Z = 1, K_PREC = 2, K_HORNER = 3, K_FUSED = 4, K_INIT = 5 };
  const size_t smem = sizeof(T) * (size_t)Cfg::template smem_elems<STEP_JACOBI>();
  const size_t smem2 = sizeof(T) * (size_t)Cfg::template smem_elems<STEP_Z>();
  if (!sh.set) {
    auto occ_of = [&](auto kern, int nt, size_t sm) {
      int o = 1;
      MVTV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, nt, sm));
      return std::max(o, 1);
    };
    if (fam == CGF_RING) {
      MVTV_CUDA(cudaFuncSetAttribute(k_cg_step<T, Cfg, STEP_JACOBI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MVTV_CUDA(cudaFuncSetAttribute(k_cg_step<T, Cfg, STEP_Z>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      MVTV_CUDA(cudaFuncSetAttribute(k_cg_step<T, Cfg, STEP_PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      sh.occ[K_STEP_J] = occ_of(k_cg_step<T, Cfg, STEP_JACOBI>, Cfg::NT, smem);
      sh.occ[K_STEP_Z] = occ_of(k_cg_step<T, Cfg, STEP_Z>, Cfg::NT, smem2);
      sh.occ[K_PREC] = occ_of(k_cg_step<T, Cfg, STEP_PREC>, Cfg::NT, smem2);
    }
    if constexpr (HAS2D) {
      if (fam == CGF_STRIP2D) {
        sh.occ[K_STEP_J] = occ_of(k_cg_step2d<T, S2Step, STEP_JACOBI>, S2Step::NT, 0);
        sh.occ[K_STEP_Z] = occ_of(k_cg_step2d<T, S2Step, STEP_Z>, S2Step::NT, 0);
        sh.occ[K_PREC] = occ_of(k_cg_step2d<T, S2Prec, STEP_PREC>, S2Prec::NT, 0);
        sh.occ[K_HORNER] = occ_of(k_cg_horner2d<T, S2_HORNER_WARPS, S2_HORNER_MINB, false>, 32 * S2_HORNER_WARPS, 0);
        sh.occ[K_FUSED] = occ_of(k_cg_updprec2d<T, S2Fuse>, S2Fuse::NT, 0);
        sh.occ[K_INIT] = occ_of(k_cg_init2d<T, S2_INIT_WARPS>, 32 * S2_INIT_WARPS, 0);
      }
    }
    if constexpr (HAS3D) {
      if (fam == CGF_STRIP3D)
        with3([&](auto ts) {
          using TS = decltype(ts);
          auto occ3 = [&](auto kern, auto cfg, int mode) {
            using C3 = decltype(cfg);
            const size_t sm = C3::smem_bytes(mode, sizeof(T));
            if (sm) MVTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            return occ_of(kern, C3::NT, sm);
          };
          sh.occ[K_STEP_J] = occ3(k_cg_step3d<T, typename TS::Step, STEP_JACOBI>, typename TS::Step{}, STEP_JACOBI);
          sh.occ[K_STEP_Z] = occ3(k_cg_step3d<T, typename TS::Step, STEP_Z>, typename TS::Step{}, STEP_Z);
          sh.occ[K_PREC] = occ3(k_cg_step3d<T, typename TS::Prec, STEP_PREC>, typename TS::Prec{}, STEP_PREC);
          sh.occ[K_INIT] = occ3(k_cg_step3d<T, typename TS::Init, STEP_INIT>, typename TS::Init{}, STEP_INIT);
          sh.occ[K_HORNER] = occ3(k_cg_step3d<T, typename TS::Prec, STEP_HORNER>, typename TS::Prec{}, STEP_HORNER);
          sh.occ[K_FUSED] = occ3(k_cg_step3d<T, typename TS::Fuse, STEP_UPDPREC>, typename TS::Fuse{}, STEP_UPDPREC);
        });
    }
    sh.set = true;
  }
  unsigned tiles_step, tiles_prec, tiles_horner = 1, tiles_fused = 1;
  if (fam == CGF_STRIP2D) {
    tiles_step = (unsigned)((m0 + S2Step::TX - 1) / S2Step::TX);
    tiles_prec = (unsigned)((m0 + S2Prec::TX - 1) / S2Prec::TX);
    tiles_horner = (unsigned)((m0 + 64 * S2_HORNER_WARPS - 1) / (64 * S2_HORNER_WARPS));
    tiles_fused = (unsigned)((m0 + S2Fuse::TX - 1) / S2Fuse::TX);
  } else if (fam == CGF_STRIP3D) {
    using TS = S3Tiles;
    tiles_step = (unsigned)(((m0 + TS::Step::TX - 1) / TS::Step::TX) * ((m1 + TS::Step::TY - 1) / TS::Step::TY));
    tiles_prec = tiles_horner = (unsigned)(((m0 + TS::Prec::TX - 1) / TS::Prec::TX) * ((m1 + TS::Prec::TY - 1) / TS::Prec::TY));
    tiles_fused = (unsigned)(((m0 + TS::Fuse::TX - 1) / TS::Fuse::TX) * ((m1 + TS::Fuse::TY - 1) / TS::Fuse::TY));
  } else {
    tiles_step = tiles_prec = (unsigned)(((m0 + Cfg::TX - 1) / Cfg::TX) * ((m1 + Cfg::TY - 1) / Cfg::TY) * ((m2 + Cfg::TW - 1) / Cfg::TW));
  }
  int zc_step = 1, zc_prec = 1, zc_horner = 1, zc_fused = 1;
  const dim3 gs = chunking(tiles_step, sh.occ[deg ? K_STEP_Z : K_STEP_J], zc_step);
  const dim3 gs_prec = chunking(tiles_prec, sh.occ[K_PREC], zc_prec);
  const dim3 gs_horner = chunking(tiles_horner, sh.occ[K_HORNER], zc_horner);
  const dim3 gs_fused = chunking(tiles_fused, sh.occ[K_FUSED], zc_fused);
  const int gu = (int)std::max<long long>(1, std::min<long long>(148 * 8, (dt.Nloc + 1023) / 1024));
  const dim3 g = grid_owned();

  // ---- r = b - M theta, theta_old = theta, r.z (Jacobi), r.r, b.b
  a.seq_red = ++red_seq;
  a.seq_halo = ++halo_seq;
  prof_begin(MVTV_KC_CG_INIT);
  bool init_done = false;
  if constexpr (HAS2D) {
    if (fam == CGF_STRIP2D) {   // marching / shuffle form (cg_init2d.cuh): 199 against 369 us on 4096^2
      const long long tiles_i = ((long long)m0 + 64 * S2_INIT_WARPS - 1) / (64 * S2_INIT_WARPS);
      long long nchunk = ((long long)nsm * sh.occ[K_INIT] + tiles_i - 1) / tiles_i;
      nchunk = std::max(1ll, std::min<long long>(nchunk, std::max(1, dt.nz / 8)));
      const int zchunk_i = (int)((dt.nz + nchunk - 1) / nchunk);
      nchunk = (dt.nz + zchunk_i - 1) / zchunk_i;
      k_cg_init2d<T, S2_INIT_WARPS><<<dim3((unsigned)tiles_i, (unsigned)nchunk, 1), 32 * S2_INIT_WARPS, 0, stream>>>(dt, st, a, RedBuf{partials, counters + 1}, zchunk_i);
      init_done = true;
    }
  }
  if constexpr (HAS3D) {
    if (init3d) {   // marching / shuffle form with TMA-staged rows (STEP_INIT of cg_step3d.cuh): 2.0 against 3.6 ms on 512^3
      int zc_init = 1;
      const dim3 gi = chunking(tiles_prec, sh.occ[K_INIT], zc_init);
      with3([&](auto ts) {
        using C3 = typename decltype(ts)::Init;
        k_cg_step3d<T, C3, STEP_INIT><<<gi, C3::NT, C3::smem_bytes(STEP_INIT, sizeof(T)), stream>>>(dt, st, a, RedBuf{partials, counters + 1}, zc_init);
      });
      init_done = true;
    }
  }
  if (!init_done) k_cg_init<T, P><<<g, block, 0, stream>>>(dt, st, a, RedBuf{partials, counters + 1});
  prof_end();
  MVTV_CUDA(cudaGetLastError());
  launches += 1;
  if (fold) {
  } else if (d_peer) {
    k_cg_peer_commit_init<<<1, 1, 0, stream>>>(S, d_peer, a.seq_red);
    launches += 1;
  } else if (world > 1) {
    allreduce(raw, 3, ncclSum);
    k_cg_commit_init<<<1, 1, 0, stream>>>(S, raw);
    launches += 1;
  }

  // ---- the launches of one CG iteration
  // first preconditioner pass, stand-alone (before the first iteration on the fused path, every iteration otherwise)
  auto launch_prec_first = [&]() {
    a.w_out_scr = wsel(1);
    a.w_in_scr = 0;
    a.seq_zout = a.seq_zhalo = z_event = ++zhalo_seq;
    a.final_pass = (deg == 1) ? 1 : 0;
    a.pc0 = hc[deg - 1];
    a.pc1 = hc[deg];
    const RedBuf rbp{partials, counters + 5};
    bool done = false;
    if constexpr (HAS2D) {
      if (fam == CGF_STRIP2D) {
        if (deg == 1) k_cg_step2d<T, S2Prec, STEP_PREC><<<gs_prec, S2Prec::NT, 0, stream>>>(dt, st, a, rbp, zc_prec);
        else k_cg_horner2d<T, S2_HORNER_WARPS, S2_HORNER_MINB, true><<<gs_horner, 32 * S2_HORNER_WARPS, 0, stream>>>(dt, st, a, rbp, zc_horner);
        done = true;
      }
    }
    if constexpr (HAS3D) {
      if (fam == CGF_STRIP3D) {
        with3([&](auto ts) {
          using C3 = typename decltype(ts)::Prec;
          k_cg_step3d<T, C3, STEP_PREC><<<gs_prec, C3::NT, C3::smem_bytes(STEP_PREC, sizeof(T)), stream>>>(dt, st, a, rbp, zc_prec);
        });
        done = true;
      }
    }
    if (!done) k_cg_step<T, Cfg, STEP_PREC><<<gs_prec, Cfg::NT, smem2, stream>>>(dt, st, a, rbp, zc_prec);
    launches += 1;
  };
  // Horner passes 2 .. deg (strip kernels, one GPU)
  auto launch_horner_rest = [&]() {
    for (int j = 2; j <= deg; ++j) {
      a.w_out_scr = wsel(j);
      a.w_in_scr = wsel(j - 1);
      a.seq_zin = z_event;
      a.seq_zout = z_event = ++zhalo_seq;
      if (j == deg) {
        a.seq_red = ++red_seq;
        a.seq_rr_pending = rr_pending;
        rr_pending = 0;
      }
      a.final_pass = (j == deg) ? 1 : 0;
      a.pc0 = hc[deg - j];
      a.pc1 = 0.0;
      const RedBuf rbp{partials, counters + 5};
      if constexpr (HAS2D) {
        if (fam == CGF_STRIP2D)
          k_cg_horner2d<T, S2_HORNER_WARPS, S2_HORNER_MINB, false><<<gs_horner, 32 * S2_HORNER_WARPS, 0, stream>>>(dt, st, a, rbp, zc_horner);
      }
      if constexpr (HAS3D) {
        if (fam == CGF_STRIP3D)
          with3([&](auto ts) {
            using C3 = typename decltype(ts)::Prec;
            k_cg_step3d<T, C3, STEP_HORNER><<<gs_horner, C3::NT, C3::smem_bytes(STEP_HORNER, sizeof(T)), stream>>>(dt, st, a, rbp, zc_horner);
          });
      }
      launches += 1;
    }
  };
  auto commit_rz = [&]() {   // world > 1 (degree 1): r.z across the ranks, then the ghost planes of z
    if (fold) {
    } else if (d_peer) {
      k_cg_peer_commit_rz<<<1, 1, 0, stream>>>(S, d_peer, a.seq_red, a.rtol2);
      launches += 1;
    } else if (world > 1) {
      allreduce(raw, 1, ncclSum);
      k_cg_commit_rz<<<1, 1, 0, stream>>>(S, raw, a.rtol2);
      exchange_ghosts<T>((T *)zbuf);
      launches += 1;
    }
  };
  auto launch_step = [&]() {   // p = z + beta p (Jacobi: z = D^-1 r on the fly), q = M p, p.q
    const RedBuf rbs{partials, counters + 2};
    a.seq_zin = z_event;                                        // ghost planes of z
    if (fused && d_peer) a.seq_zout = z_event = ++zhalo_seq;    // ... and this launch exports those of q
    bool done = false;
    if constexpr (HAS2D) {
      if (fam == CGF_STRIP2D) {
        if (deg) k_cg_step2d<T, S2Step, STEP_Z><<<gs, S2Step::NT, 0, stream>>>(dt, st, a, rbs, zc_step);
        else k_cg_step2d<T, S2Step, STEP_JACOBI><<<gs, S2Step::NT, 0, stream>>>(dt, st, a, rbs, zc_step);
        done = true;
      }
    }
    if constexpr (HAS3D) {
      if (fam == CGF_STRIP3D) {
        with3([&](auto ts) {
          using C3 = typename decltype(ts)::Step;
          if (deg) k_cg_step3d<T, C3, STEP_Z><<<gs, C3::NT, C3::smem_bytes(STEP_Z, sizeof(T)), stream>>>(dt, st, a, rbs, zc_step);
          else k_cg_step3d<T, C3, STEP_JACOBI><<<gs, C3::NT, C3::smem_bytes(STEP_JACOBI, sizeof(T)), stream>>>(dt, st, a, rbs, zc_step);
        });
        done = true;
      }
    }
    if (!done) {
      if (deg) k_cg_step<T, Cfg, STEP_Z><<<gs, Cfg::NT, smem2, stream>>>(dt, st, a, rbs, zc_step);
      else k_cg_step<T, Cfg, STEP_JACOBI><<<gs, Cfg::NT, smem, stream>>>(dt, st, a, rbs, zc_step);
    }
    launches += 1;
    if (fold) {
    } else if (d_peer) {
      k_cg_peer_commit_pq<<<1, 1, 0, stream>>>(S, d_peer, a.seq_red, a.rtol2);
      launches += 1;
    } else if (world > 1) {
      allreduce(raw, 1, ncclSum);
      k_cg_commit_pq<<<1, 1, 0, stream>>>(S, raw, a.rtol2);
      launches += 1;
    }
  };
  auto launch_update = [&]() {   // theta += alpha p, r -= alpha q, r.r (Jacobi: r.z too); fused: + first preconditioner pass
    const RedBuf rbu{partials, counters + 3};
    if (fused) {
      a.w_out_scr = wsel(1);
      a.w_in_scr = 0;
      a.seq_zin = z_event;                        // ghost planes of q
      a.seq_zout = z_event = ++zhalo_seq;
      a.defer_rr = (d_peer && deg >= 2 && tune_defer_rr) ? 1 : 0;
      rr_pending = a.defer_rr ? a.seq_red : 0;
      a.final_pass = (deg == 1) ? 1 : 0;
      a.pc0 = hc[deg - 1];
      a.pc1 = hc[deg];
      if constexpr (HAS2D) {
        if (fam == CGF_STRIP2D) k_cg_updprec2d<T, S2Fuse><<<gs_fused, S2Fuse::NT, 0, stream>>>(dt, st, a, rbu, zc_fused);
      }
      if constexpr (HAS3D) {
        if (fam == CGF_STRIP3D)
          with3([&](auto ts) {
            using C3 = typename decltype(ts)::Fuse;
            k_cg_step3d<T, C3, STEP_UPDPREC><<<gs_fused, C3::NT, C3::smem_bytes(STEP_UPDPREC, sizeof(T)), stream>>>(dt, st, a, rbu, zc_fused);
          });
      }
      launches += 1;
      return;
    }
    k_cg_update<T><<<gu, 256, 0, stream>>>(a, dt.plane, dt.Nloc, rbu);
    launches += 1;
    if (fold) {
    } else if (d_peer) {
      if (deg) k_cg_peer_commit_update_prec<<<1, 1, 0, stream>>>(S, d_peer, a.seq_red, a.rtol2);
      else k_cg_peer_commit_update<<<1, 1, 0, stream>>>(S, d_peer, a.seq_red, a.rtol2);
      launches += 1;
    } else if (world > 1) {
      allreduce(raw, 2, ncclSum);
      if (deg) k_cg_commit_update_prec<<<1, 1, 0, stream>>>(S, raw, a.rtol2);
      else k_cg_commit_update<<<1, 1, 0, stream>>>(S, raw, a.rtol2);
      launches += 1;
    }
  };

  bool first_prec_done = false;
  int launched = 0;
  int batch = std::max(2, std::min(last_cg_iters, 256));
  const int follow = std::max(2, std::min(batch / 4 + 1, 16));   // size of the follow-up batches
  double iters = 0;
  for (;;) {
    for (int k = 0; k < batch; ++k) {
      if (world > 1 && !d_peer) exchange_ghosts<T>((T *)r);
      if (deg && !(fused && first_prec_done)) {  // z = P(D^-1 M) D^-1 r and r.z (fused path: only before the first iteration)
        first_prec_done = true;
        a.seq_red = ++red_seq;       // of the pass that reduces r.z (the last one)
        prof_begin(MVTV_KC_CG_PREC);
        launch_prec_first();
        launch_horner_rest();
        prof_end();
        commit_rz();
      }
      a.seq_red = ++red_seq;       // a.seq_halo: the version the last producer of r posted
      prof_begin(MVTV_KC_CG_STEP);
      launch_step();
      prof_end();
      a.seq_red = ++red_seq;
      if (!fused) a.seq_halo = ++halo_seq;   // k_cg_update posts r's ghost planes; the fused update keeps them up to date locally
      prof_begin(MVTV_KC_CG_UPDATE);
      launch_update();
      prof_end();
      if (fused && deg >= 2) {
        prof_begin(MVTV_KC_CG_PREC);
        launch_horner_rest();
        prof_end();
      }
    }
    MVTV_CUDA(cudaGetLastError());
    launched += batch;
    read_scalars<T>(S, CS_N);
    iters = h_scal[CS_ITERS];
    const int cur = ((int)iters) & 1;
    const bool done = h_scal[2 * cur + 1] <= a.rtol2 * h_scal[CS_BB];
    if (done) break;
    if (!(h_scal[CS_BB] == h_scal[CS_BB]) || !(h_scal[2 * cur + 1] == h_scal[2 * cur + 1])) {
      status = MVTV_ERR_INNER_SOLVE;  // NaN
      break;
    }
    if (launched >= maxit) {
      // fp64: a real failure.  fp32: the recursive residual can stagnate above cg_rtol; accept theta.
      if (sizeof(T) == 8) status = MVTV_ERR_INNER_SOLVE;
      break;
    }
    // the first batch is sized from the previous x-update; a solve that needs a few iterations more gets small follow-up batches
    // (launches after convergence are no-ops, but each still costs a launch)
    batch = std::max(1, std::min(follow, maxit - launched));
  }
  inner += (long long)iters;
  last_cg_iters = (int)iters + 1;
  return (int)iters;
}


// lam_max_pinv: a truncated CG on D^T D started from Oty, then max|D x| (cpp-code/utils.cpp:354-404), or the CGNR
// variant times five (rcpp-code/MultivarTV/src/utils.cpp:306-355).  One-off setup: host-driven loop, generic kernels.
template <typename T, int P>
int mvtv_plan::lambda_max_t(int mode, double *lam_out, int *iters_out) {
  use_device();
  MVTV_REQUIRE(have_points, "mvtv_lambda_max: call mvtv_plan_set_points first");
  const long long pl = dt.plane, nl = dt.Nloc;
  const size_t vb = (size_t)dt.usz * esz();
  T *X = (T *)xold, *R = (T *)r, *Pv = (T *)pbuf[0], *AP = (T *)q, *Z0 = (T *)pbuf[1], *Dv = (T *)v1, *B = (T *)oty;
  const dim3 g = grid_owned();
  const int g1 = std::min<int>(grid1d(nl), (int)nblocks);
  MVTV_CUDA(cudaMemsetAsync(Z0, 0, vb, stream));
  auto applyK = [&](T *in, T *out) {  // out = D^T D in  (3^P-point clamped stencil, diag(c) = 0)
    exchange_ghosts<T>(in);
    k_apply_M<T, P><<<g, block, 0, stream>>>(dt, st, in, Z0, 1.0, out);
    launches += 1;
  };
  auto dot = [&](const T *a_, const T *b_) {
    k_dot<T><<<g1, 256, 0, stream>>>(a_, b_, pl, nl, RedBuf{partials, counters + 4}, zr);
    launches += 1;
    allreduce(zr, 1, ncclSum);
    read_scalars<T>(zr, 1);
    return h_scal[0];
  };
  auto axpby = [&](T *out, double a_, const T *x_, double b_, const T *y_) {
    k_axpby<T><<<g1, 256, 0, stream>>>(out, a_, x_, b_, y_, pl, nl);
    launches += 1;
  };
  int iter = 0;
  if (mode == MVTV_MODE_CPP || mode == MVTV_MODE_PY) {
    // x.fill(mean(b)); r = b - A*x; p = r     (utils.cpp:356-358)
    double sumb;
    {
      // mean(b) over ALL vertices
      MVTV_CUDA(cudaMemsetAsync(AP, 0, vb, stream));
      k_fill<T><<<grid1d(dt.usz), 256, 0, stream>>>(AP, dt.usz, (T)1.0);
      sumb = dot(B, AP);
    }
    const double meanb = sumb / (double)N;
    k_fill<T><<<grid1d(dt.usz), 256, 0, stream>>>(X, dt.usz, (T)meanb);
    applyK(X, AP);
    axpby(R, 1.0, B, -1.0, AP);
    axpby(Pv, 1.0, R, 0.0, R);
    double rsold = dot(R, R), rsnew = rsold + 1.0;
    const int MAXIT = (N < 400) ? 500 : 100;                  // utils.cpp:364-370
    while (sqrt(rsnew) >= 0.01) {                             // utils.cpp:371
      applyK(Pv, AP);
      const double alpha = rsold / dot(Pv, AP);
      axpby(X, 1.0, X, alpha, Pv);
      axpby(R, 1.0, R, -alpha, AP);
      rsnew = dot(R, R);
      iter += 1;
      if (iter == MAXIT) break;                               // "Reached max iter!" utils.cpp:378-381
      axpby(Pv, 1.0, R, rsnew / rsold, Pv);
      rsold = rsnew;
    }
  } else {
    // rcpp CGNR on A = D^T D (rcpp utils.cpp:306-340): x = 0; d = b; r = A^T d; p = r; t = A p
    MVTV_CUDA(cudaMemsetAsync(X, 0, vb, stream));
    axpby(Dv, 1.0, B, 0.0, B);
    applyK(Dv, R);
    axpby(Pv, 1.0, R, 0.0, R);
    const double rsold0 = sqrt(dot(R, R));
    double rsold = rsold0 * rsold0, rsnew = rsold + 1.0;
    applyK(Pv, AP);
    const int MAXIT = N < 2000 ? (int)N : 2000;
    while (sqrt(rsnew) >= 0.0001 * rsold0) {
      const double alpha = rsold / dot(AP, AP);
      axpby(X, 1.0, X, alpha, Pv);
      axpby(Dv, 1.0, Dv, -alpha, AP);
      applyK(Dv, R);
      rsnew = dot(R, R);
      iter += 1;
      if (iter == MAXIT) break;
      axpby(Pv, 1.0, R, rsnew / rsold, Pv);
      applyK(Pv, AP);
      rsold = rsnew;
    }
  }
  exchange_ghosts<T>(X);
  k_maxabs_D<T><<<g, block, 0, stream>>>(dt, bt, X, RedBuf{partials, counters + 4}, zr);
  launches += 1;
  allreduce(zr, 1, ncclMax);
  read_scalars<T>(zr, 1);
  MVTV_CUDA(cudaGetLastError());
  *lam_out = (mode == MVTV_MODE_RCPP) ? 5.0 * h_scal[0] : h_scal[0];   // rcpp utils.cpp:351-355
  if (iters_out) *iters_out = iter;
  dinv_rho = NAN;  // scratch vectors were borrowed from the CG loop
  return MVTV_OK;
}

template <typename T>
int mvtv_plan::solve_t(const mvtv_solve_params &prm, const double *theta_init, double *u_inout,
                       double *theta_out, double *fitted_out, mvtv_solve_result &res) {
  use_device();
  MVTV_REQUIRE(have_points, "mvtv_solve: call mvtv_plan_set_points first");
  const int mode = prm.mode;
  MVTV_REQUIRE(mode == MVTV_MODE_CPP || mode == MVTV_MODE_RCPP || mode == MVTV_MODE_PY, "bad mode");
  const double lambda = prm.lambda;
  MVTV_REQUIRE(lambda > 0.0, "lambda must be > 0");
  MVTV_REQUIRE(!(world > 1 && u_inout), "u_inout is not supported with world > 1 (use MVTV_WARM_U_FROM_PLAN)");
  const double tol = (prm.tol > 0.0) ? prm.tol : (mode == MVTV_MODE_RCPP ? 1e-4 : 1e-3);
  const int max_counter = prm.max_counter > 0 ? prm.max_counter
                          : (mode == MVTV_MODE_CPP ? 2000 : (mode == MVTV_MODE_RCPP ? 3000 : 5000));
  // delta-scaled blocks (the mbs() driver's operators, cpp-code/utils.cpp:258-267) make the system badly conditioned at
  // empty vertices: 1e-13 leaves ~2e-8 in theta there, 1e-14 keeps the 1e-9 parity bar
  bool scaled = false;
  for (int b = 0; b < bt.K; ++b) scaled = scaled || (bt.scale[b] != 1.0);
  const double cg_rtol = prm.cg_rtol > 0.0 ? prm.cg_rtol : (dtype == MVTV_F64 ? (scaled ? 1e-14 : 1e-13) : 1e-5);
  const int cg_maxit = prm.cg_maxit > 0 ? prm.cg_maxit : (dtype == MVTV_F64 ? 20000 : 1000);
  if (prm.precond < MVTV_PRECOND_JACOBI || prm.precond > MVTV_PRECOND_CHEB4)
    throw Error(MVTV_ERR_UNSUPPORTED, "precond must be one of MVTV_PRECOND_JACOBI, _CHEB1 .. _CHEB4, _AUTO");
  const long long launches0 = launches;
  const long long nvec = dt.usz;
  T *th = (T *)theta;

  // ---- initial state: cpp :92-108 ; rcpp :98-109 ; py :54-65 -----------------------------------
  if (!((prm.flags & MVTV_WARM_THETA_FROM_PLAN) && have_theta_state)) {
    if (theta_init) {
      double *stg = stage(sizeof(double) * (size_t)dt.Nloc);
      MVTV_CUDA(cudaMemcpyAsync(stg, theta_init, sizeof(double) * (size_t)dt.Nloc, cudaMemcpyHostToDevice, stream));
      k_import<T><<<grid1d(dt.Nloc), 256, 0, stream>>>(th, stg, dt.plane, dt.Nloc);
    } else {
      k_fill<T><<<grid1d(nvec), 256, 0, stream>>>(th, nvec, (T)mean_y);
    }
    launches += 1;
  }
  double rho, rhoM = (prm.rho_matrix0 == prm.rho_matrix0) ? prm.rho_matrix0 : lambda;
  if (mode == MVTV_MODE_RCPP) {
    if ((prm.flags & MVTV_WARM_U_FROM_PLAN) && have_u_state) {
      // keep u[ucur], uscale
    } else if (u_inout) {
      double *stg = stage(sizeof(double) * (size_t)rt.R);
      MVTV_CUDA(cudaMemcpyAsync(stg, u_inout, sizeof(double) * (size_t)rt.R, cudaMemcpyHostToDevice, stream));
      k_u_convert<T><<<dim3(grid1d(dt.Nloc), bt.K), 256, 0, stream>>>(dt, rt, (T *)u[ucur], stg, 1.0, 1);
      uscale = 1.0;
      launches += 1;
    } else {
      MVTV_CUDA(cudaMemsetAsync(u[ucur], 0, (size_t)nvec * bt.K * esz(), stream));
      uscale = 1.0;
    }
    rho = (prm.rho_init == prm.rho_init) ? prm.rho_init : lambda / 5.0;
  } else {
    k_fill<T><<<grid1d(nvec * bt.K), 256, 0, stream>>>((T *)u[ucur], nvec * bt.K, (T)(1.0 / lambda));
    launches += 1;
    uscale = 1.0;
    rho = (mode == MVTV_MODE_CPP) ? (double)(int)lambda : lambda;  // int rho = lambda;  cpp :108
  }
  MVTV_CUDA(cudaGetLastError());
  exchange_ghosts<T>(th);

  // first loop test of the |dtheta| modes: thetaold = mean(y) - 0.1 (cpp :103-104) / - 1 (py :63)
  double dmax = 0.0;
  if (mode != MVTV_MODE_RCPP) {
    const double c0 = mean_y - (mode == MVTV_MODE_CPP ? 0.1 : 1.0);
    k_maxabs_const<T><<<std::min<int>(grid1d(dt.Nloc), (int)nblocks), 256, 0, stream>>>(
        th, dt.plane, dt.Nloc, c0, RedBuf{partials, counters + 4}, zr);
    launches += 1;
    allreduce(zr, 1, ncclMax);
    read_scalars<T>(zr, 1);
    dmax = h_scal[0];
  }

  const int skip = prm.timing_skip_passes > 0 ? prm.timing_skip_passes : 0;   // untimed warm-up passes of a benchmark call
  long long launches_t0 = launches, inner_t0 = 0;
  int passes_t0 = 0;
  if (skip == 0) MVTV_CUDA(cudaEventRecord(ev0, stream));
  // alpha = D*theta (cpp :100): only D^T alpha and D^T u are needed, computed by the init pass
  launch_zu<T>(0.0, uscale, mode, 1, false);
  // scale still to be applied to the stored D^T u when b is formed: the initial pass above has folded a pending rescale of a
  // warm-started u into it; after a regular pass D^T u_new is stored unscaled and adapt_step's factor is pending
  double v2scale = 1.0;

  const bool trace_passes = getenv("MVTV_TRACE") != nullptr && rank == 0;
  int counter = 1, passes = 0, status = MVTV_OK;
  double dual_norm = 1.0, primal_norm = 1.0, eps_dual = tol, eps_primal = tol;  // rcpp :108-109
  double r_norm = NAN, s_norm = NAN;
  long long inner = 0;
  const double sqrtN = sqrt((double)N), sqrtR = sqrt((double)rt.R);
  for (;;) {
    if (mode == MVTV_MODE_RCPP) {
      if (!(dual_norm > eps_dual || primal_norm > eps_primal)) break;  // rcpp :110
    } else {
      if (!(dmax > tol)) break;  // cpp :113 any(abs(theta-thetaold) > TOL)
    }
    if (prm.max_passes > 0 && passes >= prm.max_passes) break;
    if (skip > 0 && passes == skip) {
      MVTV_CUDA(cudaEventRecord(ev0, stream));
      launches_t0 = launches;
      inner_t0 = inner;
      passes_t0 = passes;
    }
    // b = Oty + rho*Dt*(alpha+u) ; theta = spsolve(sp_crosses, b)     cpp :115-116 / rcpp :112-113
    int cgst = MVTV_OK;
    // AUTO: the polynomial pays off once plain Jacobi-PCG needs more than ~24 iterations (degree d needs ~0.86 (d+1) times
    // fewer iterations for d more stencil passes).  The estimate is the Jacobi-equivalent count of the previous x-updates on
    // this plan with a memory of one half per pass, so that one easy solve (e.g. the first pass of a warm-started call)
    // does not send the next, ordinary one back to Jacobi
    const double jac_last = last_cg_prec ? 0.86 * (last_cg_prec + 1) * last_cg_iters : (double)last_cg_iters;
    jac_estimate = std::max(jac_last, 0.5 * jac_estimate);
    const double jac_equiv = jac_estimate;
    int prec;
    switch (prm.precond) {
      case MVTV_PRECOND_JACOBI: prec = 0; break;
      case MVTV_PRECOND_CHEB1: prec = 1; break;
      case MVTV_PRECOND_CHEB2: prec = 2; break;
      case MVTV_PRECOND_CHEB3: prec = 3; break;
      case MVTV_PRECOND_CHEB4: prec = 4; break;
      default: prec = (jac_equiv > 24.0) ? auto_degree : 0; break;
    }
    prec = std::min(prec, max_degree);
    last_cg_prec = prec;
    switch (dt.P) {
      case 2: cg_solve<T, 2>(rho, v2scale, rhoM, cg_rtol, cg_maxit, prec, inner, cgst); break;
      case 3: cg_solve<T, 3>(rho, v2scale, rhoM, cg_rtol, cg_maxit, prec, inner, cgst); break;
      case 4: cg_solve<T, 4>(rho, v2scale, rhoM, cg_rtol, cg_maxit, prec, inner, cgst); break;
      default: throw Error(MVTV_ERR_UNSUPPORTED, "p must be 1..4");
    }
    if (trace_passes)
      fprintf(stderr, "[mvtv] pass %d: rho %.6g matrix scalar %.6g degree %d, %d CG iterations (status %d)\n", passes + 1, rho, rhoM, prec,
              last_cg_iters - 1, cgst);
    if (cgst != MVTV_OK) { status = cgst; break; }
    exchange_ghosts<T>(th);
    // alpha, residuals, u update: cpp :117-120 / rcpp :114-117
    const double kappa = (rho != 0.0) ? lambda / rho : INFINITY;
    launch_zu<T>(kappa, uscale, mode, 0, true);
    ucur ^= 1;
    uscale = 1.0;
    read_scalars<T>(zr, ZR_N);
    r_norm = sqrt(h_scal[ZR_R2]);
    s_norm = fabs(rho) * sqrt(h_scal[ZR_S2]);
    dmax = h_scal[ZR_DMAX];
    ++passes;
    v2scale = uscale;
    if (mode == MVTV_MODE_PY) continue;  // no residuals, no adaptation; counter never moves (py :65-76)
    if (mode == MVTV_MODE_CPP) {
      counter += 1;
      if (counter > max_counter) { status = MVTV_ERR_NOT_CONVERGED; break; }  // cpp :122-124
      double rho_next = rho;
      if (r_norm > 20 * s_norm) { rho_next = 20 * rho; uscale = 0.05; }        // cpp :75-79
      else if (s_norm > 20 * r_norm) { rho_next = 0.1 * rho; uscale = 10.0; }  // cpp :80-83
      rho = (double)(int)rho_next;                                             // cpp :126 (int rho)
    } else {
      dual_norm = s_norm;                                                       // rcpp :119
      primal_norm = r_norm;                                                     // rcpp :120
      eps_dual = tol * (sqrtN + sqrt(h_scal[ZR_DTU2]));                         // rcpp :121
      eps_primal = tol * (sqrtR + std::max(sqrt(h_scal[ZR_DTH2]), sqrt(h_scal[ZR_AL2])));  // rcpp :122
      if (r_norm > 10 * s_norm) { rho = 2.0 * rho; uscale = 1.0 / 2.0; }        // rcpp :82-85
      else if (s_norm > 10 * r_norm) { rho = 1.0 / 2.0 * rho; uscale = 2.0; }   // rcpp :86-89
      rhoM = rho;                                                               // rcpp :126
      counter += 1;
      if (counter > max_counter) { status = MVTV_ERR_NOT_CONVERGED; break; }    // rcpp :129-132
    }
    v2scale = uscale;
  }
  MVTV_CUDA(cudaEventRecord(ev1, stream));
  MVTV_CUDA(cudaStreamSynchronize(stream));
  prof_flush();
  float ms = 0.f;
  if (skip == 0 || passes > skip) MVTV_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  have_theta_state = true;
  have_u_state = true;
  rho_state = rho;

  // ---- outputs: theta, fitted = O*theta (cpp :65-66), u, rho ------------------------------------
  if (theta_out) {
    double *stg = stage(sizeof(double) * (size_t)dt.Nloc);
    k_export<T><<<grid1d(dt.Nloc), 256, 0, stream>>>(stg, th, dt.plane, dt.Nloc);
    MVTV_CUDA(cudaMemcpyAsync(theta_out, stg, sizeof(double) * (size_t)dt.Nloc, cudaMemcpyDeviceToHost, stream));
    MVTV_CUDA(cudaStreamSynchronize(stream));
    launches += 1;
  }
  if (fitted_out) {
    double *stg = stage(sizeof(double) * (size_t)n);
    launch_gather<T>(n, vid, th, dt.plane, dt.z0, dt.nz, stg, stream);
    MVTV_CUDA(cudaMemcpyAsync(fitted_out, stg, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, stream));
    MVTV_CUDA(cudaStreamSynchronize(stream));
    launches += 1;
  }
  if (u_inout) {
    double *stg = stage(sizeof(double) * (size_t)rt.R);
    k_u_convert<T><<<dim3(grid1d(dt.Nloc), bt.K), 256, 0, stream>>>(dt, rt, (T *)u[ucur], stg, uscale, 0);
    MVTV_CUDA(cudaMemcpyAsync(u_inout, stg, sizeof(double) * (size_t)rt.R, cudaMemcpyDeviceToHost, stream));
    MVTV_CUDA(cudaStreamSynchronize(stream));
    launches += 1;
  }
  MVTV_CUDA(cudaGetLastError());
  res.counter = counter;
  res.passes = passes;
  res.status = status;
  res.rho = rho;
  res.r_norm = r_norm;
  res.s_norm = s_norm;
  res.max_dtheta = dmax;
  res.inner_iters = inner;
  res.device_seconds = (double)ms * 1e-3;
  res.kernel_launches = launches - launches0;
  res.timed_passes = passes - passes_t0;
  res.timed_inner_iters = inner - inner_t0;
  res.timed_kernel_launches = launches - launches_t0;
  return status;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
template <typename F>
static int guarded(F &&f) {
  try {
    return f();
  } catch (const Error &e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return MVTV_ERR_INVALID;
  }
}

extern "C" {

int mvtv_abi_version(void) { return MVTV_ABI_VERSION; }
const char *mvtv_last_error(void) { return g_last_error.c_str(); }

int mvtv_device_count(int *count) {
  return guarded([&] {
    MVTV_REQUIRE(count, "null count");
    *count = 0;
    MVTV_CUDA(cudaGetDeviceCount(count));
    return MVTV_OK;
  });
}

int mvtv_nccl_unique_id(void *out128) {
  return guarded([&] {
    MVTV_REQUIRE(out128, "null argument");
    g_nccl.load();
    ncclUniqueId id;
    MVTV_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return MVTV_OK;
  });
}

int mvtv_host_alloc(void **out, uint64_t bytes) {
  return guarded([&] {
    MVTV_REQUIRE(out, "null argument");
    *out = nullptr;
    MVTV_REQUIRE(bytes > 0, "bytes must be > 0");
    MVTV_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return MVTV_OK;
  });
}

int mvtv_host_free(void *ptr) {
  return guarded([&] {
    if (ptr) MVTV_CUDA(cudaFreeHost(ptr));
    return MVTV_OK;
  });
}

int mvtv_plan_create(mvtv_plan **out, const mvtv_plan_desc *d) {
  return guarded([&] {
    MVTV_REQUIRE(out && d, "null argument");
    *out = nullptr;
    MVTV_REQUIRE(d->struct_size == (int32_t)sizeof(mvtv_plan_desc), "mvtv_plan_desc size mismatch");
    MVTV_REQUIRE(d->p >= 1 && d->p <= MVTV_MAXP, "p must be 1..4");
    MVTV_REQUIRE(d->dtype == MVTV_F64 || d->dtype == MVTV_F32, "dtype must be 64 or 32");
    MVTV_REQUIRE(d->variant == MVTV_VARIANT_REFERENCE || d->variant == MVTV_VARIANT_INTENDED, "bad variant");
    MVTV_REQUIRE(d->world >= 1 && d->rank >= 0 && d->rank < d->world, "bad rank/world");
    int ndev = 0;
    MVTV_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) throw Error(MVTV_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    std::unique_ptr<mvtv_plan> pl(new mvtv_plan());
    if (d->device >= 0) pl->device = d->device;
    else MVTV_CUDA(cudaGetDevice(&pl->device));
    MVTV_REQUIRE(pl->device < ndev, "device ordinal out of range");
    pl->dtype = d->dtype;
    pl->variant = d->variant;
    pl->rank = d->rank;
    pl->world = d->world;
    pl->build_tables(*d);
    pl->allocate();
    if (d->world > 1) {
      MVTV_REQUIRE(d->nccl_unique_id, "world > 1 needs nccl_unique_id");
      g_nccl.load();
      ncclUniqueId id;
      memcpy(&id, d->nccl_unique_id, sizeof(id));
      MVTV_NCCL(g_nccl.CommInitRank(&pl->comm, d->world, id, d->rank));
      pl->setup_peer();
    }
    pl->finalize_selection();
    *out = pl.release();
    return MVTV_OK;
  });
}

int mvtv_plan_destroy(mvtv_plan *plan) {
  return guarded([&] {
    delete plan;
    return MVTV_OK;
  });
}

int mvtv_plan_info(const mvtv_plan *plan, int64_t *N, int64_t *R, int64_t *z0, int64_t *nz) {
  return guarded([&] {
    MVTV_REQUIRE(plan, "null plan");
    if (N) *N = plan->N;
    if (R) *R = plan->rt.R;
    if (z0) *z0 = plan->dt.z0;
    if (nz) *nz = plan->dt.nz;
    return MVTV_OK;
  });
}

int mvtv_plan_describe(const mvtv_plan *plan, char *buf, int64_t cap) {
  return guarded([&] {
    MVTV_REQUIRE(plan && buf && cap > 0, "null argument");
    const char *fam = plan->cg_family == CGF_STRIP2D ? "k_cg_step2d" : (plan->cg_family == CGF_STRIP3D ? "k_cg_step3d" : "k_cg_step");
    char tmp[640];
    snprintf(tmp, sizeof(tmp),
             "{\"p\": %d, \"dtype\": %d, \"world\": %d, \"zu\": \"%s\", \"cg_step\": \"%s\", \"cg_prec\": \"%s\", "
             "\"cg_prec_words\": %d, \"fused_update\": %d, \"max_degree\": %d, \"auto_degree\": %d, \"last_degree\": %d, "
             "\"collectives\": \"%s\", \"fold_commit\": %d}",
             plan->p, plan->dtype, plan->world, plan->zu_variant >= 0 ? "k_zu_march" : "k_zu", fam, fam,
             plan->cg_family == CGF_RING ? 4 : 3, plan->fused_update ? 1 : 0, plan->max_degree, plan->auto_degree,
             plan->last_cg_prec, plan->world == 1 ? "none" : (plan->d_peer ? "peer" : "nccl"),
             (plan->d_peer && plan->fold_commit) ? 1 : 0);
    MVTV_REQUIRE((int64_t)strlen(tmp) < cap, "buffer too small");
    strcpy(buf, tmp);
    return MVTV_OK;
  });
}

int mvtv_plan_profile(mvtv_plan *plan, int enable) {
  return guarded([&] {
    MVTV_REQUIRE(plan, "null plan");
    plan->prof_on = enable != 0;
    for (int k = 0; k < MVTV_KC_N; ++k) { plan->prof_ms[k] = 0.0; plan->prof_cnt[k] = 0; }
    return MVTV_OK;
  });
}

int mvtv_plan_get_profile(mvtv_plan *plan, double *ms, int64_t *count) {
  return guarded([&] {
    MVTV_REQUIRE(plan && ms && count, "null argument");
    for (int k = 0; k < MVTV_KC_N; ++k) { ms[k] = plan->prof_ms[k]; count[k] = plan->prof_cnt[k]; }
    return MVTV_OK;
  });
}

int mvtv_plan_set_points_dev(mvtv_plan *plan, int64_t n, const double *data_dev, const double *y_dev,
                             const double *axes_dev) {
  return guarded([&] {
    MVTV_REQUIRE(plan && axes_dev && (n == 0 || (data_dev && y_dev)), "null argument");
    if (plan->dtype == MVTV_F64) plan->set_points_t<double>(n, data_dev, 1, n, y_dev, axes_dev);
    else plan->set_points_t<float>(n, data_dev, 1, n, y_dev, axes_dev);
    return MVTV_OK;
  });
}

int mvtv_plan_set_points(mvtv_plan *plan, int64_t n, const double *data, const double *y, const double *axes) {
  return mvtv_plan_set_points_strided(plan, n, data, 1, n, y, axes);
}

int mvtv_plan_set_points_strided(mvtv_plan *plan, int64_t n, const double *data, int64_t ld_point, int64_t ld_axis,
                                 const double *y, const double *axes) {
  return guarded([&] {
    MVTV_REQUIRE(plan && axes && (n == 0 || (data && y)), "null argument");
    MVTV_REQUIRE(n >= (plan->world > 1 ? 0 : 1), "n must be >= 1 (world > 1: a rank may hold no points)");
    MVTV_REQUIRE((ld_point == 1 && ld_axis == n) || (ld_point == plan->p && ld_axis == 1),
                 "data must be dense column-major (1, n) or row-major (p, 1)");
    plan->use_device();
    long long na = 0;
    for (int a = 0; a < plan->p; ++a) na += plan->dt.m[a];
    const size_t nd = (size_t)n * plan->p, total = nd + (size_t)n + (size_t)na;
    const bool trace = getenv("MVTV_TRACE") != nullptr;   // wall-clock phases of the setup on stderr
    const auto t0 = std::chrono::steady_clock::now();
    double *buf = (double *)mvtv_plan::grow(plan->in_buf, plan->in_bytes, sizeof(double) * total);
    const auto t1 = std::chrono::steady_clock::now();
    if (n > 0) {
      MVTV_CUDA(cudaMemcpyAsync(buf, data, sizeof(double) * nd, cudaMemcpyHostToDevice, plan->stream));
      MVTV_CUDA(cudaMemcpyAsync(buf + nd, y, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, plan->stream));
    }
    MVTV_CUDA(cudaMemcpyAsync(buf + nd + n, axes, sizeof(double) * (size_t)na, cudaMemcpyHostToDevice, plan->stream));
    if (trace) MVTV_CUDA(cudaStreamSynchronize(plan->stream));
    const auto t2 = std::chrono::steady_clock::now();
    if (plan->dtype == MVTV_F64) plan->set_points_t<double>(n, buf, ld_point, ld_axis, buf + nd, buf + nd + n);
    else plan->set_points_t<float>(n, buf, ld_point, ld_axis, buf + nd, buf + nd + n);
    if (trace) {
      const auto t3 = std::chrono::steady_clock::now();
      auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
      };
      fprintf(stderr, "[mvtv] set_points n=%lld: scratch %.2f ms, h2d %.2f ms (%.1f GB/s), bin+sort+reduce %.2f ms\n",
              (long long)n, ms(t0, t1), ms(t1, t2), sizeof(double) * total / (ms(t1, t2) * 1e6), ms(t2, t3));
    }
    return MVTV_OK;
  });
}

int mvtv_plan_get_cache(mvtv_plan *plan, double *Oty, double *counts, int64_t *vertex_of_point) {
  return guarded([&] {
    MVTV_REQUIRE(plan && plan->have_points, "no points set");
    plan->use_device();
    const long long nl = plan->dt.Nloc;
    double *stg = plan->stage(sizeof(double) * (size_t)nl);
    for (int k = 0; k < 2; ++k) {
      double *dst = k == 0 ? Oty : counts;
      if (!dst) continue;
      void *src = k == 0 ? plan->oty : plan->cnt;
      if (plan->dtype == MVTV_F64)
        k_export<double><<<mvtv_plan::grid1d(nl), 256, 0, plan->stream>>>(stg, (const double *)src, plan->dt.plane, nl);
      else
        k_export<float><<<mvtv_plan::grid1d(nl), 256, 0, plan->stream>>>(stg, (const float *)src, plan->dt.plane, nl);
      MVTV_CUDA(cudaMemcpyAsync(dst, stg, sizeof(double) * (size_t)nl, cudaMemcpyDeviceToHost, plan->stream));
      MVTV_CUDA(cudaStreamSynchronize(plan->stream));
    }
    if (vertex_of_point) {
      MVTV_CUDA(cudaMemcpyAsync(vertex_of_point, plan->vid, sizeof(long long) * (size_t)plan->n,
                                cudaMemcpyDeviceToHost, plan->stream));
      MVTV_CUDA(cudaStreamSynchronize(plan->stream));
    }
    return MVTV_OK;
  });
}

int mvtv_solve(mvtv_plan *plan, const mvtv_solve_params *prm, const double *theta_init, double *u_inout,
               double *theta_out, double *fitted_out, mvtv_solve_result *res) {
  return guarded([&] {
    MVTV_REQUIRE(plan && prm && res, "null argument");
    MVTV_REQUIRE(prm->struct_size == (int32_t)sizeof(mvtv_solve_params), "mvtv_solve_params size mismatch");
    memset(res, 0, sizeof(*res));
    int st;
    if (plan->dtype == MVTV_F64) st = plan->solve_t<double>(*prm, theta_init, u_inout, theta_out, fitted_out, *res);
    else st = plan->solve_t<float>(*prm, theta_init, u_inout, theta_out, fitted_out, *res);
    if (st == MVTV_ERR_NOT_CONVERGED) set_last_error("Failed to converge!");
    if (st == MVTV_ERR_INNER_SOLVE) set_last_error("x-update CG did not reach cg_rtol within cg_maxit");
    return st;
  });
}

int mvtv_solve_path(mvtv_plan *plan, const mvtv_solve_params *prm, int32_t n_lambda, const double *lambdas,
                    const double *ftrue, double *mses_out, int32_t *counters_out, double *rhos_out, double *thetas_out,
                    double *theta_best_out, double *fitted_best_out, int32_t *best_index_out, mvtv_solve_result *total) {
  return guarded([&] {
    MVTV_REQUIRE(plan && prm && lambdas && ftrue && mses_out && total, "null argument");
    MVTV_REQUIRE(prm->struct_size == (int32_t)sizeof(mvtv_solve_params), "mvtv_solve_params size mismatch");
    MVTV_REQUIRE(n_lambda >= 1, "n_lambda must be >= 1");
    MVTV_REQUIRE(plan->have_points, "call mvtv_plan_set_points first");
    plan->use_device();
    cudaStream_t s = plan->stream;
    const long long n = plan->n, nl = plan->dt.Nloc;
    double *d_target = nullptr;
    MVTV_CUDA(cudaMalloc(&d_target, sizeof(double) * (size_t)n));
    memset(total, 0, sizeof(*total));
    int rc = MVTV_OK;
    try {
      MVTV_CUDA(cudaMemcpyAsync(d_target, ftrue, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
      mvtv_solve_params p = *prm;
      double rho_carry = (p.rho_init == p.rho_init) ? p.rho_init : lambdas[0] / 5.0;  // rcpp solvers.cpp:209
      int best = 0;
      double best_mse = INFINITY;
      std::vector<double> theta_tmp;
      for (int i = 0; i < n_lambda; ++i) {
        p.lambda = lambdas[i];
        if (p.mode == MVTV_MODE_RCPP) {
          // rcpp mbs_path (solvers.cpp:212-220): theta, u, rho carried; cached matrix = crossO + rho_init*crossD
          p.rho_init = rho_carry;
          p.rho_matrix0 = rho_carry;
          p.flags = (i == 0) ? (prm->flags & ~(MVTV_WARM_THETA_FROM_PLAN | MVTV_WARM_U_FROM_PLAN))
                             : (MVTV_WARM_THETA_FROM_PLAN | MVTV_WARM_U_FROM_PLAN);
        } else {
          // cpp mbs_path (solvers.cpp:207-214): theta carried, sp_crosses = crossO + lambda*crossD, u and rho restart
          p.rho_matrix0 = lambdas[i];
          p.flags = (i == 0) ? (prm->flags & ~MVTV_WARM_THETA_FROM_PLAN) : MVTV_WARM_THETA_FROM_PLAN;
        }
        mvtv_solve_result r;
        memset(&r, 0, sizeof(r));
        double *th_out = thetas_out ? thetas_out + (size_t)i * nl : nullptr;
        int st;
        if (plan->dtype == MVTV_F64) st = plan->solve_t<double>(p, nullptr, nullptr, th_out, nullptr, r);
        else st = plan->solve_t<float>(p, nullptr, nullptr, th_out, nullptr, r);
        if (st == MVTV_ERR_NOT_CONVERGED && p.mode != MVTV_MODE_RCPP) {  // cpp: the throw leaves mbs_path
          set_last_error("Failed to converge!");
          rc = st;
          if (counters_out) counters_out[i] = r.counter;
          break;
        }
        if (st != MVTV_OK && st != MVTV_ERR_NOT_CONVERGED) { rc = st; break; }
        rho_carry = r.rho;
        // MSEs[i] = mbs_mse(tempmodel, ftrue)  (cpp :212 / rcpp :216)
        const int g = std::min<int>(mvtv_plan::grid1d(n), (int)plan->nblocks);
        if (plan->dtype == MVTV_F64)
          k_sqerr<double><<<g, 256, 0, s>>>(plan->vid, (const double *)plan->theta, d_target, n, plan->dt.plane,
                                            plan->dt.z0, RedBuf{plan->partials, plan->counters + 4}, plan->zr);
        else
          k_sqerr<float><<<g, 256, 0, s>>>(plan->vid, (const float *)plan->theta, d_target, n, plan->dt.plane,
                                           plan->dt.z0, RedBuf{plan->partials, plan->counters + 4}, plan->zr);
        MVTV_CUDA(cudaGetLastError());
        plan->launches += 1;
        double cnt = (double)n;
        if (plan->world > 1) {
          MVTV_CUDA(cudaMemcpyAsync(plan->zr + 1, &cnt, sizeof(double), cudaMemcpyHostToDevice, s));
          plan->allreduce(plan->zr, 2, ncclSum);
        }
        MVTV_CUDA(cudaMemcpyAsync(plan->h_scal, plan->zr, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
        MVTV_CUDA(cudaStreamSynchronize(s));
        if (plan->world > 1) cnt = plan->h_scal[1];
        const double m = plan->h_scal[0] / cnt;
        mses_out[i] = m;
        if (counters_out) counters_out[i] = r.counter;
        if (rhos_out) rhos_out[i] = r.rho;
        if (i == 0 || m < best_mse) {  // first instance of the lowest MSE (cpp solvers.cpp:172-175); lambda 0 is
                                       // always captured, so an all-NaN path still returns a model
          best_mse = m;
          best = i;
          if (theta_best_out || fitted_best_out) {
            theta_tmp.resize((size_t)nl);
            double *stg = plan->stage(sizeof(double) * (size_t)std::max(nl, n));
            if (plan->dtype == MVTV_F64) k_export<double><<<mvtv_plan::grid1d(nl), 256, 0, s>>>(stg, (const double *)plan->theta, plan->dt.plane, nl);
            else k_export<float><<<mvtv_plan::grid1d(nl), 256, 0, s>>>(stg, (const float *)plan->theta, plan->dt.plane, nl);
            MVTV_CUDA(cudaMemcpyAsync(theta_tmp.data(), stg, sizeof(double) * (size_t)nl, cudaMemcpyDeviceToHost, s));
            MVTV_CUDA(cudaStreamSynchronize(s));
          }
        }
        total->passes += r.passes;
        total->counter += r.counter;
        total->inner_iters += r.inner_iters;
        total->device_seconds += r.device_seconds;
        total->kernel_launches += r.kernel_launches;
        total->rho = r.rho;
        total->r_norm = r.r_norm;
        total->s_norm = r.s_norm;
        total->max_dtheta = r.max_dtheta;
        if (st == MVTV_ERR_NOT_CONVERGED) total->status = st;
      }
      if (best_index_out) *best_index_out = best;
      if (rc == MVTV_OK && theta_best_out) memcpy(theta_best_out, theta_tmp.data(), sizeof(double) * (size_t)nl);
      if (rc == MVTV_OK && fitted_best_out) {
        // fitted = O * theta_best through the predict gather on the staged copy
        double *stg = plan->stage(sizeof(double) * (size_t)(plan->dt.usz + n));
        MVTV_CUDA(cudaMemsetAsync(stg, 0, sizeof(double) * (size_t)plan->dt.usz, s));
        MVTV_CUDA(cudaMemcpyAsync(stg + plan->dt.plane, theta_tmp.data(), sizeof(double) * (size_t)nl, cudaMemcpyHostToDevice, s));
        launch_gather<double>(n, plan->vid, stg, plan->dt.plane, plan->dt.z0, plan->dt.nz, stg + plan->dt.usz, s);
        MVTV_CUDA(cudaMemcpyAsync(fitted_best_out, stg + plan->dt.usz, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
        MVTV_CUDA(cudaStreamSynchronize(s));
      }
    } catch (...) {
      cudaFree(d_target);
      throw;
    }
    MVTV_CUDA(cudaFree(d_target));
    return rc;
  });
}

int mvtv_lambda_max(mvtv_plan *plan, int mode, double *lambda_max, int32_t *cg_iters) {
  return guarded([&] {
    MVTV_REQUIRE(plan && lambda_max, "null argument");
    int it = 0;
    int rc;
#define MVTV_LM(TT)                                                          \
  switch (plan->dt.P) {                                                      \
    case 2: rc = plan->lambda_max_t<TT, 2>(mode, lambda_max, &it); break;    \
    case 3: rc = plan->lambda_max_t<TT, 3>(mode, lambda_max, &it); break;    \
    case 4: rc = plan->lambda_max_t<TT, 4>(mode, lambda_max, &it); break;    \
    default: throw Error(MVTV_ERR_UNSUPPORTED, "p must be 1..4");            \
  }
    if (plan->dtype == MVTV_F64) { MVTV_LM(double) } else { MVTV_LM(float) }
#undef MVTV_LM
    if (cg_iters) *cg_iters = it;
    return rc;
  });
}

int mvtv_predict(mvtv_plan *plan, int64_t n_new, const double *data, const double *axes, const double *theta,
                 double *fits_out) {
  return guarded([&] {
    MVTV_REQUIRE(plan && data && axes && fits_out, "null argument");
    MVTV_REQUIRE(n_new >= 1, "n_new must be >= 1");
    plan->use_device();
    cudaStream_t s = plan->stream;
    long long na = 0;
    for (int a = 0; a < plan->p; ++a) na += plan->dt.m[a];
    const long long nl = plan->dt.Nloc;
    const size_t nd = (size_t)n_new * plan->p;
    // staging: [data | axes | vid (as 8-byte) | out | theta(ghosted, double)]
    const size_t total = nd + (size_t)na + (size_t)n_new * 2 + (size_t)plan->dt.usz;
    double *buf = plan->stage(sizeof(double) * total);
    double *d_data = buf, *d_axes = buf + nd;
    long long *d_vid = (long long *)(d_axes + na);
    double *d_out = (double *)(d_vid + n_new);
    double *d_theta = d_out + n_new;
    MVTV_CUDA(cudaMemcpyAsync(d_data, data, sizeof(double) * nd, cudaMemcpyHostToDevice, s));
    MVTV_CUDA(cudaMemcpyAsync(d_axes, axes, sizeof(double) * (size_t)na, cudaMemcpyHostToDevice, s));
    launch_bin(plan->p, plan->dt, n_new, d_data, 1, n_new, d_axes, d_vid, nullptr, nullptr, s);
    if (theta) {
      MVTV_CUDA(cudaMemcpyAsync(d_theta + plan->dt.plane, theta, sizeof(double) * (size_t)nl, cudaMemcpyHostToDevice, s));
      launch_gather<double>(n_new, d_vid, d_theta, plan->dt.plane, plan->dt.z0, plan->dt.nz, d_out, s);
    } else {
      MVTV_REQUIRE(plan->have_theta_state, "theta == NULL needs a previous mvtv_solve on this plan");
      if (plan->dtype == MVTV_F64)
        launch_gather<double>(n_new, d_vid, (const double *)plan->theta, plan->dt.plane, plan->dt.z0, plan->dt.nz, d_out, s);
      else
        launch_gather<float>(n_new, d_vid, (const float *)plan->theta, plan->dt.plane, plan->dt.z0, plan->dt.nz, d_out, s);
    }
    MVTV_CUDA(cudaMemcpyAsync(fits_out, d_out, sizeof(double) * (size_t)n_new, cudaMemcpyDeviceToHost, s));
    MVTV_CUDA(cudaStreamSynchronize(s));
    plan->launches += 2;
    return MVTV_OK;
  });
}

int mvtv_apply_D(mvtv_plan *plan, const double *theta, double *out_rows) {
  return guarded([&] {
    MVTV_REQUIRE(plan && theta && out_rows, "null argument");
    MVTV_REQUIRE(plan->world == 1, "operator-level entry points are single-GPU");
    plan->use_device();
    cudaStream_t s = plan->stream;
    const size_t nv = (size_t)plan->dt.usz, R = (size_t)plan->rt.R;
    double *buf = plan->stage(sizeof(double) * (nv + R + 1));
    MVTV_CUDA(cudaMemsetAsync(buf, 0, sizeof(double) * nv, s));
    MVTV_CUDA(cudaMemcpyAsync(buf + plan->dt.plane, theta, sizeof(double) * (size_t)plan->dt.Nloc, cudaMemcpyHostToDevice, s));
    k_apply_D<<<dim3(mvtv_plan::grid1d(plan->dt.Nloc), plan->bt.K), 256, 0, s>>>(plan->dt, plan->bt, plan->rt, buf, buf + nv);
    MVTV_CUDA(cudaGetLastError());
    MVTV_CUDA(cudaMemcpyAsync(out_rows, buf + nv, sizeof(double) * R, cudaMemcpyDeviceToHost, s));
    MVTV_CUDA(cudaStreamSynchronize(s));
    plan->launches += 1;
    return MVTV_OK;
  });
}

int mvtv_apply_Dt(mvtv_plan *plan, const double *rows, double *out_vertices) {
  return guarded([&] {
    MVTV_REQUIRE(plan && rows && out_vertices, "null argument");
    MVTV_REQUIRE(plan->world == 1, "operator-level entry points are single-GPU");
    plan->use_device();
    cudaStream_t s = plan->stream;
    const size_t nv = (size_t)plan->dt.usz, R = (size_t)plan->rt.R;
    double *buf = plan->stage(sizeof(double) * (nv + R + 1));
    MVTV_CUDA(cudaMemcpyAsync(buf + nv, rows, sizeof(double) * R, cudaMemcpyHostToDevice, s));
    k_apply_Dt<<<plan->grid_owned(), 256, 0, s>>>(plan->dt, plan->bt, plan->rt, buf + nv, buf);
    MVTV_CUDA(cudaGetLastError());
    MVTV_CUDA(cudaMemcpyAsync(out_vertices, buf + plan->dt.plane, sizeof(double) * (size_t)plan->dt.Nloc,
                              cudaMemcpyDeviceToHost, s));
    MVTV_CUDA(cudaStreamSynchronize(s));
    plan->launches += 1;
    return MVTV_OK;
  });
}

int mvtv_apply_M(mvtv_plan *plan, double sc, const double *x, double *out) {
  return guarded([&] {
    MVTV_REQUIRE(plan && x && out, "null argument");
    MVTV_REQUIRE(plan->world == 1, "operator-level entry points are single-GPU");
    MVTV_REQUIRE(plan->have_points, "mvtv_apply_M needs the point counts: call mvtv_plan_set_points first");
    plan->use_device();
    cudaStream_t s = plan->stream;
    const size_t nv = (size_t)plan->dt.usz;
    double *buf = plan->stage(sizeof(double) * nv * 3);
    double *dx = buf, *dc = buf + nv, *dout = buf + 2 * nv;
    MVTV_CUDA(cudaMemsetAsync(buf, 0, sizeof(double) * nv * 3, s));
    MVTV_CUDA(cudaMemcpyAsync(dx + plan->dt.plane, x, sizeof(double) * (size_t)plan->dt.Nloc, cudaMemcpyHostToDevice, s));
    if (plan->dtype == MVTV_F64)
      MVTV_CUDA(cudaMemcpyAsync(dc, plan->cnt, sizeof(double) * nv, cudaMemcpyDeviceToDevice, s));
    else {
      // counts are small integers: widen exactly
      k_export<float><<<mvtv_plan::grid1d(plan->dt.Nloc), 256, 0, s>>>(dc + plan->dt.plane, (const float *)plan->cnt,
                                                                       plan->dt.plane, plan->dt.Nloc);
    }
    const dim3 g = plan->grid_owned();
    switch (plan->dt.P) {
      case 2: k_apply_M<double, 2><<<g, 256, 0, s>>>(plan->dt, plan->st, dx, dc, sc, dout); break;
      case 3: k_apply_M<double, 3><<<g, 256, 0, s>>>(plan->dt, plan->st, dx, dc, sc, dout); break;
      case 4: k_apply_M<double, 4><<<g, 256, 0, s>>>(plan->dt, plan->st, dx, dc, sc, dout); break;
      default: throw Error(MVTV_ERR_UNSUPPORTED, "p must be 1..4");
    }
    MVTV_CUDA(cudaGetLastError());
    MVTV_CUDA(cudaMemcpyAsync(out, dout + plan->dt.plane, sizeof(double) * (size_t)plan->dt.Nloc, cudaMemcpyDeviceToHost, s));
    MVTV_CUDA(cudaStreamSynchronize(s));
    plan->launches += 1;
    return MVTV_OK;
  });
}

int mvtv_softthresh(int64_t n, const double *z, double lam, double *out) {
  return guarded([&] {
    MVTV_REQUIRE(n >= 0 && (n == 0 || (z && out)), "bad argument");
    if (n == 0) return (int)MVTV_OK;
    double *buf = nullptr;
    MVTV_CUDA(cudaMalloc(&buf, sizeof(double) * (size_t)n * 2));
    cudaError_t e = cudaMemcpy(buf, z, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      k_softthresh<<<mvtv_plan::grid1d(n), 256>>>(buf, lam, buf + n, n);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, buf + n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    MVTV_CUDA(e);
    return (int)MVTV_OK;
  });
}

// adapt_step (cpp-code/solvers.cpp:70-88 ; rcpp solvers.cpp:77-94) as a stand-alone call: ||r||, ||s|| by the
// deterministic grid reduction, then the rescale of u.  The ADMM loop itself never calls this: there the norms come
// out of k_zu_march and the rescale is the lazy `uscale`.
int mvtv_adapt_step(int mode, int64_t n_r, const double *r, int64_t n_s, const double *s_vec, double rho, int64_t n_u,
                    const double *u, double *rho_next, double *u_next) {
  return guarded([&] {
    MVTV_REQUIRE(mode == MVTV_MODE_CPP || mode == MVTV_MODE_RCPP, "mode must be MVTV_MODE_CPP or MVTV_MODE_RCPP");
    MVTV_REQUIRE(n_r >= 1 && n_s >= 1 && n_u >= 0 && r && s_vec && rho_next && (n_u == 0 || (u && u_next)), "bad argument");
    const int g = mvtv_plan::grid1d(std::max(n_r, n_s));
    const size_t words = (size_t)n_r + (size_t)n_s + (size_t)n_u + (size_t)g + 4;
    double *buf = nullptr;
    MVTV_CUDA(cudaMalloc(&buf, sizeof(double) * words));
    double *d_r = buf, *d_s = buf + n_r, *d_u = d_s + n_s, *d_part = d_u + n_u, *d_out = d_part + g;
    unsigned *d_cnt = (unsigned *)(d_out + 2);
    double h[2] = {0.0, 0.0};
    cudaError_t e = cudaMemset(d_out, 0, sizeof(double) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_r, r, sizeof(double) * (size_t)n_r, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_s, s_vec, sizeof(double) * (size_t)n_s, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_u) e = cudaMemcpy(d_u, u, sizeof(double) * (size_t)n_u, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      k_dot<double><<<mvtv_plan::grid1d(n_r), 256>>>(d_r, d_r, 0, n_r, RedBuf{d_part, d_cnt}, d_out);
      k_dot<double><<<mvtv_plan::grid1d(n_s), 256>>>(d_s, d_s, 0, n_s, RedBuf{d_part, d_cnt}, d_out + 1);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h, d_out, sizeof(double) * 2, cudaMemcpyDeviceToHost);
    double factor = 1.0;
    if (e == cudaSuccess) {
      const double r_norm = sqrt(h[0]), s_norm = sqrt(h[1]);
      const double thr = (mode == MVTV_MODE_CPP) ? 20.0 : 10.0;
      *rho_next = rho;
      if (r_norm > thr * s_norm) {          // cpp :75-79 (x20, x0.05) ; rcpp :82-85 (x2, x1/2)
        *rho_next = (mode == MVTV_MODE_CPP) ? 20 * rho : 2.0 * rho;
        factor = (mode == MVTV_MODE_CPP) ? 0.05 : 1.0 / 2.0;
      } else if (s_norm > thr * r_norm) {   // cpp :80-83 (x0.1, x10) ; rcpp :86-89 (x1/2, x2)
        *rho_next = (mode == MVTV_MODE_CPP) ? 0.1 * rho : 1.0 / 2.0 * rho;
        factor = (mode == MVTV_MODE_CPP) ? 10.0 : 2.0;
      }
      if (n_u) {
        if (factor != 1.0) {
          k_scale<double><<<mvtv_plan::grid1d(n_u), 256>>>(d_u, n_u, factor);
          e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpy(u_next, d_u, sizeof(double) * (size_t)n_u, cudaMemcpyDeviceToHost);
      }
    }
    cudaFree(buf);
    MVTV_CUDA(e);
    return (int)MVTV_OK;
  });
}

int mvtv_nearest(int p, const int64_t *m, const double *axes, int64_t n, const double *data, int64_t *vertex_out) {
  return guarded([&] {
    MVTV_REQUIRE(p >= 1 && p <= MVTV_MAXP && m && axes && data && vertex_out && n >= 1, "bad argument");
    DimTab dt{};
    dt.P = p;
    long long na = 0, s = 1;
    for (int a = 0; a < p; ++a) {
      dt.m[a] = m[a];
      dt.stride[a] = s;
      s *= m[a];
      na += m[a];
    }
    const size_t nd = (size_t)n * p;
    double *buf = nullptr;
    MVTV_CUDA(cudaMalloc(&buf, sizeof(double) * (nd + (size_t)na + (size_t)n)));
    long long *d_vid = (long long *)(buf + nd + na);
    cudaError_t e = cudaMemcpy(buf, data, sizeof(double) * nd, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(buf + nd, axes, sizeof(double) * (size_t)na, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      try {
        launch_bin(p, dt, n, buf, 1, n, buf + nd, d_vid, nullptr, nullptr, nullptr);
      } catch (...) {
        cudaFree(buf);
        throw;
      }
      e = cudaMemcpy(vertex_out, d_vid, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost);
    }
    cudaFree(buf);
    MVTV_CUDA(e);
    return (int)MVTV_OK;
  });
}

}  // extern "C"
