// Device code of the ADMM iteration (sm_100a).  All kernels are templates on the storage/arithmetic
// type T (double | float); grid-wide reductions always accumulate in double.
//
// Kernel inventory (reference op each one replaces):
//   k_zu        fused z/u update: D*theta - u -> softthresh -> u += r, D^T(alpha), D^T(u), D^T(u_old),
//               ||r||^2 ||s||^2 ||D^T u||^2 ||D theta||^2 ||alpha||^2 max|dtheta|
//               (cpp-code/solvers.cpp:115,117-120,113 ; rcpp solvers.cpp:112,114-117,119-122)
//   k_cg_init   b = Oty + rho*D^T(alpha+u) (never stored), r = b - M theta, p = M_J^-1 r     (solvers.cpp:115-116)
//   k_cg_step   p = M_J^-1 r + beta p fused with q = (diag(c) + s*D^T D) p (3^P-point clamped stencil out of
//               shared-memory plane tiles, marching along the last axis), p.q                 (solvers.cpp:116)
//   k_cg_update theta += a p, r -= a q, r.z, r.r
#pragma once
#include "mvtv_internal.cuh"

// dynamic shared memory of a kernel (the CPU emulator of tests/cuda_emu hands out one host buffer instead)
#ifndef MVTV_DYN_SMEM
#define MVTV_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace mvtv {

// slots of the ADMM reduction
enum { ZR_R2 = 0, ZR_S2 = 1, ZR_DTU2 = 2, ZR_DTH2 = 3, ZR_AL2 = 4, ZR_DMAX = 5, ZR_N = 6, ZR_NSUM = 5 };

// device scalars of the CG loop (doubles).  {r.z, r.r} live in two slots selected by the parity of
// the number of iterations PERFORMED so far, so that once the done test fires every later launch sees
// the same (frozen) state and returns immediately.
enum {
  CS_RZ0 = 0, CS_RR0 = 1, CS_RZ1 = 2, CS_RR1 = 3,
  CS_BB = 4,     // b.b
  CS_PQ = 5,     // p.q of the current iteration
  CS_ITERS = 6,  // iterations performed
  CS_N = 8
};

template <typename T>
struct ZuArgs {
  const T *theta;       // ghosted slab
  const T *theta_prev;  // ghosted slab (theta before the x-update) or nullptr
  const T *u_old;       // K blocks of usz
  T *u_new;
  T *v1, *v2;           // D^T alpha, D^T u_new (ghosted slab, owned part written)
  double kappa;         // lambda / rho, may be +inf
  double uscale;        // lazy rescale of adapt_step: true u_old = uscale * stored u_old
  int mode;             // MVTV_MODE_*
  int init;             // 1: alpha := D*theta, u unchanged (cpp-code/solvers.cpp:100-101), nothing written to u
  double *red_out;      // ZR_N totals
};

template <typename T>
__device__ __forceinline__ T soft_threshold(T z, T kappa) {
  // cpp-code/solvers.cpp:24-29: sign(z) % max(|z| - kappa, 0)
  T mag = fabs(z) - kappa;
  mag = mag > T(0) ? mag : T(0);
  T sg = (z > T(0)) ? T(1) : ((z < T(0)) ? T(-1) : T(0));
  return sg * mag;
}

// Decode the in-plane coordinates of vertex q and return the two boundary bitmasks:
// lo_ok bit a: i_a >= 1 ; hi_ok bit a: i_a <= m_a - 2.
__device__ __forceinline__ void boundary_masks(const DimTab &dt, long long q, int zl, int &lo_ok, int &hi_ok) {
  lo_ok = 0;
  hi_ok = 0;
  unsigned rem = (unsigned)q;
  for (int a = 0; a < dt.P - 1; ++a) {
    const unsigned ma = (unsigned)dt.m[a];
    const unsigned ia = rem % ma;
    rem /= ma;
    lo_ok |= (ia >= 1u) << a;
    hi_ok |= (ia + 2u <= ma) << a;
  }
  const long long gz = dt.z0 + zl;
  lo_ok |= (gz >= 1) << (dt.P - 1);
  hi_ok |= (gz + 2 <= dt.m[dt.P - 1]) << (dt.P - 1);
}

// ---------------------------------------------------------------------------------------------
// k_zu (gather form).  One thread per vertex v of the slab (+ one plane of "ghost row" threads when a
// lower neighbour rank exists).  For every block b and every subset e of its axis set the thread
// re-derives the row owned by vertex v-e (D*theta from theta, u_old from memory), so the three
// transposed products at v need no communication between threads; the thread whose e==0 owns the
// row and is the only one that writes u_new and counts it in the norms.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_zu(const __grid_constant__ DimTab dt, const __grid_constant__ BlockTab bt, const ZuArgs<T> a,
     const RedBuf rb) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zl = (int)blockIdx.y - dt.has_lo;  // -1: ghost-row plane
  double red[ZR_N] = {0, 0, 0, 0, 0, 0};
  if (q < dt.plane) {
    int lo_ok, hi_ok;
    boundary_masks(dt, q, zl, lo_ok, hi_ok);
    const long long base = (long long)(zl + 1) * dt.plane + q;
    const bool owned = zl >= 0;
    const int lastbit = 1 << (dt.P - 1);
    const T kappa = (T)a.kappa;
    const T usc = (T)a.uscale;
    T acc1 = 0, acc2 = 0, acc3 = 0;
    for (int b = 0; b < bt.K; ++b) {
      const int S = bt.mask[b];
      if (!owned && !(S & lastbit)) continue;
      const int ns = bt.nsub[b];
      const T sc = (T)bt.scale[b];
      const T *uo_b = a.u_old + (size_t)b * dt.usz;
      T *un_b = a.u_new + (size_t)b * dt.usz;
      const int jmax = owned ? ns : 1;
      for (int j = 0; j < jmax; ++j) {
        const int e = bt.sub[b][j];
        if ((e & ~lo_ok) != 0) continue;           // v-e leaves the mesh
        if (((S & ~e) & ~hi_ok) != 0) continue;    // vertex v-e owns no row of this block
        const long long w = base - bt.off[b][j];
        T d = 0;
        for (int f = 0; f < ns; ++f) {
          const T t = a.theta[w + bt.off[b][f]];
          d += (__popc(bt.sub[b][f]) & 1) ? -t : t;
        }
        d *= sc;
        const T uo = usc * uo_b[w];
        T al, un, pr;
        if (a.init) {
          al = d; un = uo; pr = 0;
        } else {
          al = soft_threshold<T>(d - uo, kappa);   // cpp :117
          pr = al - d;                             // cpp :119
          un = uo + pr;                            // cpp :120
        }
        const T sg = (__popc(e) & 1) ? -sc : sc;
        acc1 += sg * al;
        acc2 += sg * un;
        acc3 += sg * uo;
        if (j == 0) {
          if (!a.init) un_b[w] = un;
          if (owned) {
            red[ZR_R2] += (double)pr * (double)pr;
            red[ZR_DTH2] += (double)d * (double)d;
            red[ZR_AL2] += (double)al * (double)al;
          }
        }
      }
    }
    if (owned) {
      a.v1[base] = acc1;
      a.v2[base] = acc2;
      // dual residual without its rho factor: cpp :118 D^T(alpha + u_old) ; rcpp :117 D^T(u_new - u_old)
      const double sv = (a.mode == MVTV_MODE_RCPP) ? (double)acc2 - (double)acc3 : (double)acc1 + (double)acc3;
      red[ZR_S2] = sv * sv;
      red[ZR_DTU2] = (double)acc2 * (double)acc2;
      if (a.theta_prev) red[ZR_DMAX] = fabs((double)a.theta[base] - (double)a.theta_prev[base]);
    }
  }
  double *out = a.red_out;
  grid_reduce<ZR_N, ZR_NSUM>(red, rb, [out](const double (&res)[ZR_N]) {
#pragma unroll
    for (int k = 0; k < ZR_N; ++k) out[k] = res[k];
  });
}

// ---------------------------------------------------------------------------------------------
// Clamped 3^P-point stencil, fully unrolled at compile time.
// ---------------------------------------------------------------------------------------------
template <typename T, int A>
struct StencilRec {
  static __device__ __forceinline__ void run(const T *__restrict__ x, long long idx, int cidx,
                                             const long long *om, const long long *op,
                                             const StencilTab &st, T &acc) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const long long o = (d == 0) ? om[A] : ((d == 2) ? op[A] : 0);
      StencilRec<T, A - 1>::run(x, idx + o, cidx * 3 + d, om, op, st, acc);
    }
  }
};
template <typename T>
struct StencilRec<T, -1> {
  static __device__ __forceinline__ void run(const T *__restrict__ x, long long idx, int cidx,
                                             const long long *, const long long *, const StencilTab &st,
                                             T &acc) {
    acc += (T)st.coef[cidx] * x[idx];
  }
};

// per-vertex neighbour offsets with Neumann clamping; returns the boundary class for diag(K)
template <int P>
__device__ __forceinline__ int neighbour_offsets(const DimTab &dt, long long q, int zl, long long (&om)[P],
                                                 long long (&op)[P]) {
  int cls = 0;
  unsigned rem = (unsigned)q;
#pragma unroll
  for (int a = 0; a < P - 1; ++a) {
    const unsigned ma = (unsigned)dt.m[a];
    const unsigned ia = rem % ma;
    rem /= ma;
    om[a] = (ia >= 1u) ? -dt.stride[a] : 0;
    op[a] = (ia + 2u <= ma) ? dt.stride[a] : 0;
    cls |= ((ia == 0u) || (ia + 1u == ma)) << a;
    // an axis of extent 1 has no neighbours at all: diag contribution handled by coef (L = 0)
  }
  const long long gz = dt.z0 + zl;
  om[P - 1] = (gz >= 1) ? -dt.plane : 0;
  op[P - 1] = (gz + 2 <= dt.m[P - 1]) ? dt.plane : 0;
  cls |= ((gz == 0) || (gz + 1 == dt.m[P - 1])) << (P - 1);
  return cls;
}

// ---------------------------------------------------------------------------------------------
// Peer-memory collectives for the CG loop (one process per GPU, buffers shared with CUDA IPC over
// NVLink/NVSwitch).  Instead of launching NCCL kernels between the compute kernels, the reducing kernel's
// last thread stores its partial sums straight into every peer's slot and raises a flag; the consumer sums
// the world's partials in rank order (bitwise identical on every rank).  The kernel that produces r also
// stores its boundary planes into the neighbours' ghost planes, so the CG iteration needs no NCCL call.
// ---------------------------------------------------------------------------------------------
#define MVTV_PEER_MAXW 8
#define MVTV_PEER_NSLOT 4
#define MVTV_PEER_NVAL 4
struct PeerTab {
  int rank, world, has_lo, has_hi;
  double *slots[MVTV_PEER_MAXW];              // peer j's reduction slots  [NSLOT][MAXW][NVAL]
  unsigned long long *flags[MVTV_PEER_MAXW];  // peer j's arrival flags    [NSLOT][MAXW]
  unsigned long long *hflag_at_prev, *hflag_at_next;      // halo-ready flags we RAISE in the neighbours' buffers
  unsigned long long *hflag_from_prev, *hflag_from_next;  // ... and the ones we WAIT on in our own buffer
  void *rghost_at_prev, *rghost_at_next;      // the neighbours' ghost planes of r that this rank fills
  void *zghost_at_prev, *zghost_at_next;      // ... and of z (polynomial preconditioner)
  unsigned long long *zflag_at_prev, *zflag_at_next, *zflag_from_prev, *zflag_from_next;
  int *error;                                 // set when a wait times out (a peer died): results become NaN
  // 3-D strip kernels (fused update, Horner passes): the neighbours' ghost planes of q and of the buffers a preconditioner
  // pass may write -- PW_Z, PW_P0 / PW_P1 (the idle direction buffer), PW_Y.  All of them signal on the z flags, with one
  // increasing event number per exchange (seq_zout of the producer = seq_zin of the consumer).
  void *qghost_at_prev, *qghost_at_next;
  void *wghost_at_prev[4], *wghost_at_next[4];
};
enum { PW_Z = 0, PW_P0 = 1, PW_P1 = 2, PW_Y = 3 };

#ifdef MVTV_CUDA_EMU   // tests/cuda_emu: the kernel source compiled for the CPU emulator (ranks are host threads: host atomics)
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_relaxed_sys(double *p, double v) { __atomic_store(p, &v, __ATOMIC_RELAXED); }
__device__ __forceinline__ void st_flag_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
__device__ __forceinline__ double ld_relaxed_sys(const double *p) { double v; __atomic_load(p, &v, __ATOMIC_RELAXED); return v; }
#else
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(double *p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// flag store of a release PATTERN: the caller has executed __threadfence_system() after the data stores; several flags (one
// per peer) then cost one system-scope fence instead of one per st.release
__device__ __forceinline__ void st_flag_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
#endif
// returns false on timeout (~10 s): never hang the GPU on a dead peer
__device__ __forceinline__ bool peer_spin(const unsigned long long *flag, unsigned long long seq, int *error) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq) {
    if (clock64() - t0 > 20000000000ll) {
      if (error) *error = 1;
      return false;
    }
  }
  return true;
}
// single thread: publish n partial values of reduction event `seq` to every rank (self included)
__device__ __forceinline__ void peer_post(const PeerTab &pt, unsigned long long seq, const double *v, int n) {
  const int slot = (int)(seq % MVTV_PEER_NSLOT);
  for (int j = 0; j < pt.world; ++j)
    for (int k = 0; k < n; ++k) st_relaxed_sys(pt.slots[j] + (slot * MVTV_PEER_MAXW + pt.rank) * MVTV_PEER_NVAL + k, v[k]);
  __threadfence_system();
  for (int j = 0; j < pt.world; ++j) st_flag_sys(pt.flags[j] + slot * MVTV_PEER_MAXW + pt.rank, seq);
}
// single thread: wait for every rank's partials of event `seq`, add them in rank order
__device__ __forceinline__ void peer_wait_sum(const PeerTab &pt, unsigned long long seq, double *out, int n) {
  const int slot = (int)(seq % MVTV_PEER_NSLOT);
  for (int k = 0; k < n; ++k) out[k] = 0.0;
  bool ok = true;
  for (int j = 0; j < pt.world; ++j) {
    ok = peer_spin(pt.flags[pt.rank] + slot * MVTV_PEER_MAXW + j, seq, pt.error) && ok;
    for (int k = 0; k < n; ++k) out[k] += ld_relaxed_sys(pt.slots[pt.rank] + (slot * MVTV_PEER_MAXW + j) * MVTV_PEER_NVAL + k);
  }
  if (!ok)
    for (int k = 0; k < n; ++k) out[k] = nan("");
}

template <typename T>
struct CgArgs {
  T *x;            // theta (ghosted slab), updated in place
  T *xold;         // copy of theta before the x-update (for max|dtheta|)
  T *r, *q;        // ghosted slabs
  T *pbuf[2];      // search direction, ping-pong by the parity of the iterations performed
  const T *c;      // diag(O^T O) = points per vertex
  const T *dinv;   // 1 / (c + rhoM * diag(D^T D)): Jacobi preconditioner
  const T *oty, *v1, *v2;
  double *S;       // CS_* scalars
  double *raw;     // multi-GPU: reductions land here, are all-reduced, then committed by k_cg_commit_*;
                   // nullptr on one GPU (the reducing kernel's last thread commits directly)
  double rho;      // multiplies D^T(alpha+u) in b
  double uscale;   // lazy rescale of u: b = Oty + rho*(v1 + uscale*v2)
  double rhoM;     // scalar of the system matrix diag(c) + rhoM * D^T D
  double rtol2;    // cg_rtol^2
  const PeerTab *peer;          // non-null: peer-memory collectives (world > 1, CUDA IPC available)
  unsigned long long seq_red;   // reduction event id this launch posts / commits
  unsigned long long seq_halo;  // version of r's ghost planes this launch posts (init, update) or needs (step / prec)
  unsigned long long seq_zhalo; // version of z's ghost planes (polynomial preconditioner: prec posts, step needs)
  T *z;            // polynomial preconditioner: z = P(D^-1 M) D^-1 r (ghosted slab)
  double pc0, pc1; // first Horner pass: w_1 = pc0*z0 + pc1*D^-1 M z0,  z0 = D^-1 r  (degree 1: w_1 = z)
  int prec;        // 0: Jacobi (z = D^-1 r formed on the fly), d >= 1: degree-d Chebyshev polynomial in D^-1 M
  // degree >= 2 (Horner form, one stencil pass per degree): pass outputs rotate over the z buffer, X = the direction buffer
  // that is idle between two step kernels -- pbuf[parity ^ 1] once the update has advanced the iteration count, pbuf[parity]
  // inside the fused update (cg_wsel) -- and the extra buffer y, ending in z: [z], [X, z], [X, y, z], [z, X, y, z] for degree
  // 1..4.  q stays readable, and on several GPUs no buffer is rewritten before the pass after next, i.e. not before the
  // neighbour that reads its ghost planes has finished (its completion flag is what the pass in between waited for).
  int w_out_scr = 0;         // buffer this pass writes: 0 = z, 1 = X, 2 = y
  int w_in_scr = 0;          // buffer pass k >= 2 reads w_{k-1} from:  w_k = D^-1 M w_{k-1} + pc0*z0
  int final_pass = 1;        // the pass that produces z also reduces r.z
  T *y = nullptr;
  unsigned long long seq_zin = 0, seq_zout = 0;   // k_cg_step3d: z-flag event this launch needs / posts (see PeerTab)
  // several GPUs, degree >= 2: the fused update only POSTS its partial r.r (event seq_red) and advances the iteration count;
  // the last Horner pass of the iteration collects it together with its own r.z (seq_rr_pending = that event, 0 = none),
  // so an iteration has two world-wide rendezvous (p.q ; r.z + r.r) whatever the degree
  unsigned long long seq_rr_pending = 0;
  int defer_rr = 0;
  // fused update + first preconditioner pass (k_cg_updprec*): r is updated OUT OF PLACE, the buffer holding the
  // current residual is selected by the parity of the iterations performed (r2 == nullptr: r is updated in place)
  T *r2 = nullptr;
  int fold;        // peer path (default; MVTV_FOLD_COMMIT=0 turns it off): the reducing kernel's last thread also waits for the world's
                   // partials and commits the scalars, instead of a separate one-thread k_cg_peer_commit_* launch
};

// buffer `sel` of a preconditioner pass (see CgArgs::w_out_scr); before_commit: called by the kernel that advances the
// iteration count.  widx: the same choice as an index into PeerTab::wghost_at_*.
template <typename T>
__device__ __forceinline__ T *cg_wsel(const CgArgs<T> &a, int sel, int iters, bool before_commit) {
  return sel == 0 ? a.z : (sel == 2 ? a.y : a.pbuf[before_commit ? (iters & 1) : ((iters & 1) ^ 1)]);
}
__device__ __forceinline__ int cg_widx(int sel, int iters, bool before_commit) {
  return sel == 0 ? PW_Z : (sel == 2 ? PW_Y : (PW_P0 + (before_commit ? (iters & 1) : ((iters & 1) ^ 1))));
}
// the buffer holding the current residual (see CgArgs::r2)
template <typename T>
__device__ __forceinline__ T *cg_rcur(const CgArgs<T> &a, int iters) { return (a.r2 && (iters & 1)) ? a.r2 : a.r; }

__device__ __forceinline__ bool cg_done(const double *S, double rtol2) {
  const int cur = ((int)S[CS_ITERS]) & 1;
  return S[2 * cur + 1] <= rtol2 * S[CS_BB];
}
__device__ __forceinline__ void cg_commit_init(double *S, const double *r3) {
  S[CS_RZ0] = r3[0];
  S[CS_RR0] = r3[1];
  S[CS_BB] = r3[2];
  S[CS_RZ1] = 0.0;
  S[CS_RR1] = 0.0;
  S[CS_PQ] = 1.0;
  S[CS_ITERS] = 0.0;
}
__device__ __forceinline__ void cg_commit_update(double *S, const double *r2) {
  const int nxt = (((int)S[CS_ITERS]) & 1) ^ 1;
  S[2 * nxt] = r2[0];
  S[2 * nxt + 1] = r2[1];
  S[CS_ITERS] += 1.0;
}

// polynomial-preconditioner variants: r.z comes from the prec kernel (slot of the CURRENT parity), the update only
// delivers r.r and advances the iteration count
__device__ __forceinline__ void cg_commit_rz(double *S, const double *r1) {
  const int cur = ((int)S[CS_ITERS]) & 1;
  S[2 * cur] = r1[0];
}
__device__ __forceinline__ void cg_commit_update_prec(double *S, const double *r1) {
  const int nxt = (((int)S[CS_ITERS]) & 1) ^ 1;
  S[2 * nxt + 1] = r1[0];
  S[CS_ITERS] += 1.0;
}

// dinv = 1 / (c + rhoM*diag(K)) on every plane that holds real data (ghost planes included)
template <typename T, int P>
__global__ void __launch_bounds__(256)
k_make_dinv(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const T *c, double rhoM,
            T *dinv) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zl = (int)blockIdx.y - 1;  // -1 .. nz
  if (q >= dt.plane) return;
  if ((zl < 0 && !dt.has_lo) || (zl >= dt.nz && !dt.has_hi)) return;
  long long om[P], op[P];
  const int cls = neighbour_offsets<P>(dt, q, zl, om, op);
  const long long base = (long long)(zl + 1) * dt.plane + q;
  dinv[base] = T(1) / (c[base] + (T)rhoM * (T)st.diagK[cls]);
}

// r = b - M theta with b = Oty + rho*(D^T alpha + uscale*D^T u) (b is never stored); r.z, r.r, b.b
template <typename T, int P>
__global__ void __launch_bounds__(256)
k_cg_init(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
          const RedBuf rb) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zl = blockIdx.y;
  double red[3] = {0, 0, 0};
  if (q < dt.plane) {
    long long om[P], op[P];
    neighbour_offsets<P>(dt, q, zl, om, op);
    const long long base = (long long)(zl + 1) * dt.plane + q;
    T kx = 0;
    StencilRec<T, P - 1>::run(a.x, base, 0, om, op, st, kx);
    const T xv = a.x[base];
    const T bv = a.oty[base] + (T)a.rho * (a.v1[base] + (T)a.uscale * a.v2[base]);
    const T rv = bv - (a.c[base] * xv + (T)a.rhoM * kx);
    const T zv = rv * a.dinv[base];
    a.r[base] = rv;
    a.xold[base] = xv;
    if (a.peer) {  // fill the neighbours' ghost planes of r
      if (zl == 0 && a.peer->has_lo) { ((T *)a.peer->rghost_at_prev)[q] = rv; __threadfence_system(); }
      if (zl == dt.nz - 1 && a.peer->has_hi) { ((T *)a.peer->rghost_at_next)[q] = rv; __threadfence_system(); }
    }
    red[0] = (double)rv * (double)zv;
    red[1] = (double)rv * (double)rv;
    red[2] = (double)bv * (double)bv;
  }
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

// plain M*x for the ABI-level mvtv_apply_M
template <typename T, int P>
__global__ void __launch_bounds__(256)
k_apply_M(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const T *x, const T *c,
          double s, T *out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zl = blockIdx.y;
  if (q < dt.plane) {
    long long om[P], op[P];
    neighbour_offsets<P>(dt, q, zl, om, op);
    const long long base = (long long)(zl + 1) * dt.plane + q;
    T kx = 0;
    StencilRec<T, P - 1>::run(x, base, 0, om, op, st, kx);
    out[base] = c[base] * x[base] + (T)s * kx;
  }
}

// ---------------------------------------------------------------------------------------------
// k_cg_step: fused CG direction update + SpMV.
//   p_new = dinv.*r + beta*p_old        (never re-read from HBM: built tile by tile in shared memory)
//   q     = (diag(c) + rhoM*K) p_new    (3^P-point clamped stencil)
//   p.q
// A CTA owns an in-plane tile (TX x TY x TW over axes 0..Q-1, Q = P-1) and marches along the last axis
// over a chunk of planes.
//   * staging: the raw tiles (+ 1-deep halo, clamped at the mesh boundary) of r, dinv and p_old are
//     copied global -> shared with cp.async (LDGSTS) into a DEPTH-stage ring, DEPTH-1 planes ahead of
//     the consumer, so HBM latency is covered by the ring and not by occupancy;
//   * combine: p_new of the plane is formed once in a shared tile and the owned part written back;
//   * stencil: reuse along the marching axis happens in registers: when plane z arrives every thread
//     adds its contribution to the three output planes z-1, z, z+1 (three running accumulator sets) and
//     retires plane z-1.  In-plane a thread owns RY consecutive outputs along axis 1, so a (RY+2) x 3
//     window feeds RY outputs and lanes run along axis 0 (conflict-free LDS).
// Algorithmic traffic: read r, dinv, p_old, c ; write p_new, q  (6 N).
// ---------------------------------------------------------------------------------------------
template <int Q_, int TXT_, int XO_, int TY_, int RY_, int TW_, int DEPTH_>
struct StepCfg {
  static constexpr int Q = Q_;
  static constexpr int TXT = TXT_;          // threads along axis 0
  static constexpr int XO = XO_;            // strided outputs per thread along axis 0
  static constexpr int TX = TXT_ * XO_, TY = TY_, TW = TW_, RY = RY_;
  static constexpr int NT = TXT_ * (TY_ / RY_) * TW_;
  static constexpr int EX = TX + 2, EY = (Q_ >= 2 ? TY_ + 2 : 1), EW = (Q_ >= 3 ? TW_ + 2 : 1);
  static constexpr int TE = EX * EY * EW;   // plane tile with halo
  static constexpr int NE = (TE + NT - 1) / NT;
  static constexpr int DEPTH = DEPTH_;      // cp.async ring stages
  static constexpr int SMEM_ELEMS = (3 * DEPTH_ + 1) * TE;     // STEP_JACOBI stages three arrays per plane
  static constexpr int SMEM_ELEMS2 = (2 * DEPTH_ + 1) * TE;    // STEP_Z / STEP_PREC stage two
  // 2-D meshes (Q == 1) run best with the three-array ring stride for every variant (fewer, longer-lived CTAs);
  // 3-D / 4-D use the compact two-array ring for STEP_Z / STEP_PREC to fit one more CTA per SM
  template <int MODE>
  __host__ __device__ static constexpr int narr() { return (MODE == 0 || Q_ == 1) ? 3 : 2; }
  template <int MODE>
  __host__ __device__ static constexpr int smem_elems() { return (narr<MODE>() * DEPTH_ + 1) * TE; }
  static_assert(TY_ % RY_ == 0, "RY must divide TY");
  static_assert(Q_ >= 2 || (TY_ == 1 && RY_ == 1), "Q=1 has no axis 1");
  static_assert(Q_ >= 3 || TW_ == 1, "Q<3 has no axis 2");
};

#ifdef MVTV_CUDA_EMU   // CPU emulator: the copy completes at once, groups are no-ops
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src) { memcpy(smem_dst, gmem_src, BYTES); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

// ---- TMA bulk copies global -> shared, tracked by an mbarrier (cp.async.bulk: SASS UBLKCP; no tensor map is needed for
// contiguous rows).  One thread arms the barrier with the byte count of a stage, any thread issues the copies, the consumers
// wait for the stage's phase parity.  16-byte aligned addresses, sizes in multiples of 16 bytes.
#ifdef MVTV_CUDA_EMU   // CPU emulator: the copy completes at once, the barrier is never waited for
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *) { memcpy(smem_dst, gmem_src, bytes); }
__device__ __forceinline__ void mbar_wait(unsigned long long *, unsigned) {}
#else
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {   // make the initialised barriers visible to the async proxy
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}
#endif

// MODE: what the staged arrays are and what leaves the kernel
//   STEP_JACOBI  stage r, dinv, p_old : p = dinv.*r + beta*p_old ; writes p, q = M p ; reduces p.q
//   STEP_Z       stage z, p_old       : p = z + beta*p_old       ; writes p, q = M p ; reduces p.q
//   STEP_PREC    stage r, dinv        : z0 = dinv.*r ; writes z = pc0*z0 + pc1*dinv.*(M z0) ; reduces r.z
enum { STEP_JACOBI = 0, STEP_Z = 1, STEP_PREC = 2, STEP_HORNER = 3, STEP_UPDPREC = 4, STEP_INIT = 5 };   // the last three: cg_step3d.cuh

template <typename T, typename Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::NT, (Cfg::NT <= 256 ? (MODE != STEP_PREC ? 4 : 3) : 1))
k_cg_step(const __grid_constant__ DimTab dt, const __grid_constant__ StencilTab st, const CgArgs<T> a,
          const RedBuf rb, const int zchunk) {
  if (cg_done(a.S, a.rtol2)) return;
  constexpr int Q = Cfg::Q, TXT = Cfg::TXT, XO = Cfg::XO, TX = Cfg::TX, TY = Cfg::TY, TW = Cfg::TW, RY = Cfg::RY;
  constexpr int NT = Cfg::NT, EX = Cfg::EX, EY = Cfg::EY, TE = Cfg::TE, NE = Cfg::NE, DEPTH = Cfg::DEPTH;
  constexpr int NDY = (Q >= 2) ? 3 : 1, NDW = (Q >= 3) ? 3 : 1;   // in-plane stencil extents beyond axis 0
  constexpr int PW = 3 * NDY * NDW;                                 // stencil points per plane
  MVTV_DYN_SMEM(smem_raw);
  constexpr int NARR = Cfg::template narr<MODE>();     // staged arrays per plane (ring stride)
  constexpr int SLOT_B = 1;                              // dinv (JACOBI, PREC)
  constexpr int SLOT_C = (MODE == STEP_Z && NARR == 2) ? 1 : 2;  // p_old (JACOBI, Z)
  T *ring = reinterpret_cast<T *>(smem_raw);   // [DEPTH][NARR][TE]: r|z, dinv, p_old
  T *pn = ring + NARR * DEPTH * TE;            // [TE]: p_new of the plane being consumed

  const int tid = threadIdx.x;
  const int it = (int)a.S[CS_ITERS];
  const int cur = it & 1;
  const bool first = (MODE == STEP_PREC) ? true : (it == 0);   // "first": no p_old term
  const T beta = first ? T(0) : (T)(a.S[2 * cur] / a.S[2 * (cur ^ 1)]);
  const T *__restrict__ p_in = a.pbuf[cur];
  T *__restrict__ p_out = a.pbuf[cur ^ 1];
  const T *__restrict__ rr = (MODE == STEP_Z) ? a.z : cg_rcur(a, it);     // first staged array
  T *__restrict__ zo = cg_wsel(a, a.w_out_scr, it, false);     // STEP_PREC output
  const T *__restrict__ dinv = a.dinv;
  const T rhoM = (T)a.rhoM;

  // in-plane tile origin
  const int m0 = (int)dt.m[0];
  const int m1 = (Q >= 2) ? (int)dt.m[1] : 1;
  const int m2 = (Q >= 3) ? (int)dt.m[2] : 1;
  const int ntx = (m0 + TX - 1) / TX;
  const int nty = (m1 + TY - 1) / TY;
  int bid = blockIdx.x;
  const int bx = bid % ntx;
  bid /= ntx;
  const int by = bid % nty;
  const int bw = bid / nty;
  const int x0 = bx * TX, y0 = by * TY, w0 = bw * TW;

  // staging bookkeeping: clamped in-plane source offset of each tile element this thread stages, and whether the
  // element is an interior one inside the mesh (those are the vertices whose p_new this CTA writes back)
  int src[NE];
  bool wr[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) {
    const int e = tid + k * NT;
    int ex = e % EX, ey = (e / EX) % EY, ew = e / (EX * EY);
    int gx = x0 + ex - 1, gy = (Q >= 2) ? y0 + ey - 1 : 0, gw = (Q >= 3) ? w0 + ew - 1 : 0;
    bool interior = (ex >= 1 && ex <= TX && gx < m0);
    if (Q >= 2) interior = interior && (ey >= 1 && ey <= TY && gy < m1);
    if (Q >= 3) interior = interior && (ew >= 1 && ew <= TW && gw < m2);
    gx = min(max(gx, 0), m0 - 1);
    gy = min(max(gy, 0), m1 - 1);
    gw = min(max(gw, 0), m2 - 1);
    src[k] = gx + m0 * (gy + m1 * gw);
    wr[k] = interior && (e < TE);
  }

  const int zc0 = blockIdx.y * zchunk;
  const int zc1 = min(zc0 + zchunk, dt.nz);
  const int zlo = dt.has_lo ? -1 : 0;          // lowest / highest local plane that holds real data
  const int zhi = dt.has_hi ? dt.nz : dt.nz - 1;
  const int zfirst = zc0 - 1, zlast = zc1;     // planes consumed by this CTA
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

  // issue the async copies of plane zz (clamped) into ring stage (zz - zfirst) % DEPTH; always commits a group
  auto stage_plane = [&](int zz) {
    if (zz <= zlast) {
      const int zs = min(max(zz, zlo), zhi);
      const long long pb = (long long)(zs + 1) * dt.plane;
      T *dst = ring + (size_t)((zz - zfirst) % DEPTH) * NARR * TE;
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        const int e = tid + k * NT;
        if (e < TE) {
          const long long idx = pb + src[k];
          cp_async<sizeof(T)>(dst + e, rr + idx);
          if (MODE != STEP_Z) cp_async<sizeof(T)>(dst + SLOT_B * TE + e, dinv + idx);
          if (!first) cp_async<sizeof(T)>(dst + SLOT_C * TE + e, p_in + idx);
        }
      }
    }
    cp_async_commit();
  };

  // this thread's outputs: x = tx + i*TXT (i < XO), y = yb*RY + j (j < RY), w = tw
  const int tx = tid % TXT;
  const int yb = (tid / TXT) % (TY / RY);
  const int tw = tid / (TXT * (TY / RY));
  T A0[XO][RY], A1[XO][RY], A2[XO][RY], pcp[XO][RY];
  T cq0[XO][RY], cq1[XO][RY], cq2[XO][RY];   // diag(c) of the planes retiring now / next / after next
  T rcp[XO][RY], dcp[XO][RY], rcn[XO][RY], dcn[XO][RY];  // STEP_PREC: r and dinv at the outputs of the planes retiring next / after
  long long oidx[XO][RY];                    // in-plane offset of each output, -1 outside the mesh
#pragma unroll
  for (int i = 0; i < XO; ++i)
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      A0[i][j] = A1[i][j] = A2[i][j] = pcp[i][j] = T(0);
      cq0[i][j] = cq1[i][j] = cq2[i][j] = T(0);
      rcp[i][j] = dcp[i][j] = rcn[i][j] = dcn[i][j] = T(0);
      const int gx = x0 + tx + i * TXT, gy = y0 + yb * RY + j, gw = w0 + tw;
      oidx[i][j] = (gx < m0 && gy < m1 && gw < m2) ? gx + (long long)m0 * (gy + (long long)m1 * gw) : -1;
    }
  // register prefetch of diag(c), two planes ahead of its use in the retire step
  auto fetch_c = [&](int z, T (&dst)[XO][RY]) {
    if (z >= zc0 && z < zc1) {
      const long long pb = (long long)(z + 1) * dt.plane;
#pragma unroll
      for (int i = 0; i < XO; ++i)
#pragma unroll
        for (int j = 0; j < RY; ++j)
          if (oidx[i][j] >= 0) dst[i][j] = a.c[pb + oidx[i][j]];
    }
  };

  double red[1] = {0.0};
#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d) stage_plane(zfirst + d);
  fetch_c(zc0, cq1);
  for (int zz = zfirst; zz <= zlast; ++zz) {
    stage_plane(zz + DEPTH - 1);
    fetch_c(zz + 1, cq2);
    cp_async_wait<DEPTH - 1>();   // this thread's copies of plane zz have landed
    __syncthreads();              // ... and everybody else's; also: all threads are done reading pn (plane zz-1)
    // ---- combine: p_new of plane zz, written back for planes this CTA owns (plus the ghost planes, which the
    // first / last chunk keep up to date redundantly so p never needs a halo exchange)
    {
      const int zs = min(max(zz, zlo), zhi);
      const bool own = (zz == zs) && ((zz >= zc0 && zz < zc1) || (zz < 0 && zc0 == 0) || (zz >= dt.nz && zc1 == dt.nz));
      const long long pb = (long long)(zs + 1) * dt.plane;
      const T *stg = ring + (size_t)((zz - zfirst) % DEPTH) * NARR * TE;
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        const int e = tid + k * NT;
        if (e < TE) {
          T v = (MODE == STEP_Z) ? stg[e] : stg[SLOT_B * TE + e] * stg[e];
          if (!first) v += beta * stg[SLOT_C * TE + e];
          pn[e] = v;
          if (MODE != STEP_PREC && own && wr[k]) p_out[pb + src[k]] = v;
        }
      }
      if (MODE == STEP_PREC) {
        // r and dinv of plane zz at this thread's outputs, straight from the ring stage -- read BEFORE the barrier:
        // after it, faster threads refill this stage with plane zz + DEPTH
#pragma unroll
        for (int i = 0; i < XO; ++i)
#pragma unroll
          for (int j = 0; j < RY; ++j) {
            const int e0 = (tx + i * TXT + 1) + EX * (((Q >= 2) ? yb * RY + j + 1 : 0) + EY * ((Q >= 3) ? tw + 1 : 0));
            rcn[i][j] = stg[e0];
            dcn[i][j] = stg[SLOT_B * TE + e0];
          }
      }
    }
    __syncthreads();
    // ---- stencil contributions of plane zz to output planes zz+1 (A2), zz (A1), zz-1 (A0)
    T pcc[XO][RY];
#pragma unroll
    for (int i = 0; i < XO; ++i) {
      const int ex = tx + i * TXT + 1;
#pragma unroll
      for (int dw = 0; dw < NDW; ++dw) {
        const int ew = (Q >= 3) ? tw + dw : 0;
        constexpr int NR = (Q >= 2) ? RY + 2 : 1;
        T W[NR][3];
#pragma unroll
        for (int rrow = 0; rrow < NR; ++rrow) {
          const int ey = (Q >= 2) ? yb * RY + rrow : 0;
          const T *row = pn + ex - 1 + EX * (ey + EY * ew);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) W[rrow][dx] = row[dx];
        }
#pragma unroll
        for (int j = 0; j < RY; ++j) {
          if (dw == ((Q >= 3) ? 1 : 0)) pcc[i][j] = W[(Q >= 2) ? j + 1 : 0][1];
#pragma unroll
          for (int dy = 0; dy < NDY; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const T w = W[(Q >= 2) ? j + dy : 0][dx];
              const int ci = dx + 3 * (((Q >= 2) ? dy : 0) + NDY * ((Q >= 3) ? dw : 0));  // in-plane coefficient index
              A2[i][j] += (T)st.coef[ci] * w;            // plane zz is the z-1 neighbour of output plane zz+1
              A1[i][j] += (T)st.coef[ci + PW] * w;       // ... the centre plane of output plane zz
              A0[i][j] += (T)st.coef[ci + 2 * PW] * w;   // ... the z+1 neighbour of output plane zz-1
            }
        }
      }
    }
    // ---- retire output plane zz-1
    if (zz - 1 >= zc0) {
      const long long pb = (long long)zz * dt.plane;   // local plane zz-1 sits at (zz-1+1)*plane
#pragma unroll
      for (int i = 0; i < XO; ++i)
#pragma unroll
        for (int j = 0; j < RY; ++j)
          if (oidx[i][j] >= 0) {
            const T pv = pcp[i][j];
            const T qv = cq0[i][j] * pv + rhoM * A0[i][j];
            if (MODE == STEP_PREC) {
              const T zv = (T)a.pc0 * pv + (T)a.pc1 * (dcp[i][j] * qv);
              zo[pb + oidx[i][j]] = zv;
              if (a.peer) {  // fill the neighbours' ghost planes of z
                if (zz - 1 == 0 && dt.has_lo) { ((T *)a.peer->zghost_at_prev)[oidx[i][j]] = zv; __threadfence_system(); }
                if (zz - 1 == dt.nz - 1 && dt.has_hi) { ((T *)a.peer->zghost_at_next)[oidx[i][j]] = zv; __threadfence_system(); }
              }
              red[0] += (double)rcp[i][j] * (double)zv;
            } else {
              a.q[pb + oidx[i][j]] = qv;
              red[0] += (double)pv * (double)qv;
            }
          }
    }
    if (MODE == STEP_PREC) {
#pragma unroll
      for (int i = 0; i < XO; ++i)
#pragma unroll
        for (int j = 0; j < RY; ++j) { rcp[i][j] = rcn[i][j]; dcp[i][j] = dcn[i][j]; }
    }
#pragma unroll
    for (int i = 0; i < XO; ++i)
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        A0[i][j] = A1[i][j];
        A1[i][j] = A2[i][j];
        A2[i][j] = T(0);
        pcp[i][j] = pcc[i][j];
        cq0[i][j] = cq1[i][j];
        cq1[i][j] = cq2[i][j];
      }
  }
  cp_async_wait<0>();
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

// theta += alpha p ; r -= alpha q ; r.z ; r.r     (flat streaming kernel over the owned slab)
template <typename T>
__global__ void __launch_bounds__(256)
k_cg_update(const CgArgs<T> a, const long long plane, const long long nloc, const RedBuf rb) {
  if (cg_done(a.S, a.rtol2)) return;
  T *gprev = (a.peer && a.peer->has_lo) ? (T *)a.peer->rghost_at_prev : nullptr;
  T *gnext = (a.peer && a.peer->has_hi) ? (T *)a.peer->rghost_at_next : nullptr;
  bool stored_peer = false;
  const int cur = ((int)a.S[CS_ITERS]) & 1;
  const T alpha = (T)(a.S[2 * cur] / a.S[CS_PQ]);
  const T *__restrict__ p = a.pbuf[cur ^ 1] + plane;
  const T *__restrict__ q = a.q + plane;
  const T *__restrict__ dinv = a.dinv + plane;
  T *__restrict__ x = a.x + plane;
  T *__restrict__ r = a.r + plane;
  double red[2] = {0, 0};
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nloc; i += 4 * stride) {
    T pv[4], qv[4], rv[4], xv[4], dv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long j = i + k * stride;
      pv[k] = p[j]; qv[k] = q[j]; rv[k] = r[j]; xv[k] = x[j]; dv[k] = a.prec ? T(0) : dinv[j];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long j = i + k * stride;
      const T rn = rv[k] - alpha * qv[k];
      x[j] = xv[k] + alpha * pv[k];
      r[j] = rn;
      if (gprev && j < plane) { gprev[j] = rn; stored_peer = true; }
      if (gnext && j >= nloc - plane) { gnext[j - (nloc - plane)] = rn; stored_peer = true; }
      red[0] += (double)rn * (double)(rn * dv[k]);
      red[1] += (double)rn * (double)rn;
    }
  }
  for (; i < nloc; i += stride) {
    const T rn = r[i] - alpha * q[i];
    x[i] += alpha * p[i];
    r[i] = rn;
    if (gprev && i < plane) { gprev[i] = rn; stored_peer = true; }
    if (gnext && i >= nloc - plane) { gnext[i - (nloc - plane)] = rn; stored_peer = true; }
    if (!a.prec) red[0] += (double)rn * (double)(rn * dinv[i]);
    red[1] += (double)rn * (double)rn;
  }
  if (stored_peer) __threadfence_system();
  double *S = a.S, *raw = a.raw;
  const PeerTab *peer = a.peer;
  const unsigned long long sr = a.seq_red, sh = a.seq_halo;
  const int prec = a.prec;
  const int fold = a.fold;
  grid_reduce<2, 2>(red, rb, [S, raw, peer, sr, sh, prec, fold](const double (&res)[2]) {
    if (peer) {
      __threadfence_system();
      if (peer->has_lo) st_release_sys(peer->hflag_at_prev, sh);
      if (peer->has_hi) st_release_sys(peer->hflag_at_next, sh);
      peer_post(*peer, sr, res, 2);
      if (fold) {   // what k_cg_peer_commit_update / k_cg_peer_commit_update_prec do
        double v[2];
        peer_wait_sum(*peer, sr, v, 2);
        if (prec) cg_commit_update_prec(S, v + 1);
        else cg_commit_update(S, v);
      }
    } else if (raw) { raw[0] = res[0]; raw[1] = res[1]; }
    else if (prec) cg_commit_update_prec(S, res + 1);
    else cg_commit_update(S, res);
  });
}

// multi-GPU commits over peer memory: wait for the world's partials of the event, sum in rank order
__global__ void k_cg_peer_commit_init(double *S, const PeerTab *peer, unsigned long long seq) {
  double v[3];
  peer_wait_sum(*peer, seq, v, 3);
  cg_commit_init(S, v);
}
__global__ void k_cg_peer_commit_pq(double *S, const PeerTab *peer, unsigned long long seq, double rtol2) {
  if (cg_done(S, rtol2)) return;
  double v[1];
  peer_wait_sum(*peer, seq, v, 1);
  S[CS_PQ] = v[0];
}
__global__ void k_cg_peer_commit_update(double *S, const PeerTab *peer, unsigned long long seq, double rtol2) {
  if (cg_done(S, rtol2)) return;
  double v[2];
  peer_wait_sum(*peer, seq, v, 2);
  cg_commit_update(S, v);
}

__global__ void k_cg_peer_commit_rz(double *S, const PeerTab *peer, unsigned long long seq, double rtol2) {
  if (cg_done(S, rtol2)) return;
  double v[1];
  peer_wait_sum(*peer, seq, v, 1);
  cg_commit_rz(S, v);
}
__global__ void k_cg_peer_commit_update_prec(double *S, const PeerTab *peer, unsigned long long seq, double rtol2) {
  if (cg_done(S, rtol2)) return;
  double v[2];
  peer_wait_sum(*peer, seq, v, 2);
  cg_commit_update_prec(S, v + 1);
}
__global__ void k_cg_commit_rz(double *S, const double *raw, double rtol2) {
  if (cg_done(S, rtol2)) return;
  cg_commit_rz(S, raw);
}
__global__ void k_cg_commit_update_prec(double *S, const double *raw, double rtol2) {
  if (cg_done(S, rtol2)) return;
  cg_commit_update_prec(S, raw + 1);
}

// multi-GPU commits (after the all-reduce of `raw`)
__global__ void k_cg_commit_init(double *S, const double *raw) { cg_commit_init(S, raw); }
__global__ void k_cg_commit_pq(double *S, const double *raw, double rtol2) {
  if (cg_done(S, rtol2)) return;
  S[CS_PQ] = raw[0];
}
__global__ void k_cg_commit_update(double *S, const double *raw, double rtol2) {
  if (cg_done(S, rtol2)) return;
  cg_commit_update(S, raw);
}

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_fill(T *x, long long n, T v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}
template <typename T>
__global__ void k_scale(T *x, long long n, T s) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= s;
}
// dst (ghosted slab, type T) <- src (contiguous owned part, double) and back
template <typename T>
__global__ void k_import(T *dst_ghosted, const double *src, long long plane, long long nloc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (long long)gridDim.x * blockDim.x)
    dst_ghosted[i + plane] = (T)src[i];
}
template <typename T>
__global__ void k_export(double *dst, const T *src_ghosted, long long plane, long long nloc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (double)src_ghosted[i + plane];
}
__global__ void k_softthresh(const double *z, double lam, double *out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = soft_threshold<double>(z[i], lam);
}

// ---------------------------------------------------------------------------------------------
// ABI-boundary kernels in the reference's compact row layout (single GPU only)
// ---------------------------------------------------------------------------------------------
// direction 0: padded u (type T, times uscale) -> compact rows (double) ; 1: compact -> padded
template <typename T>
__global__ void k_u_convert(const __grid_constant__ DimTab dt, const __grid_constant__ RowTab rt, T *u_padded,
                            double *rows, double uscale, int direction) {
  const int b = blockIdx.y;
  const long long nrow = rt.rows[b];
  const int S = rt.mask[b];
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nrow; r += (long long)gridDim.x * blockDim.x) {
    long long rem = r, v = 0;
    for (int a = 0; a < dt.P; ++a) {
      const long long rd = dt.m[a] - ((S >> a) & 1);
      v += (rem % rd) * dt.stride[a];
      rem /= rd;
    }
    const size_t pi = (size_t)b * dt.usz + dt.plane + v;
    if (direction == 0) rows[rt.row_off[b] + r] = uscale * (double)u_padded[pi];
    else u_padded[pi] = (T)rows[rt.row_off[b] + r];
  }
}

// out_rows = D * theta  (compact layout, double in/out; theta ghosted slab of doubles)
__global__ void k_apply_D(const __grid_constant__ DimTab dt, const __grid_constant__ BlockTab bt,
                          const __grid_constant__ RowTab rt, const double *theta_ghosted, double *rows) {
  const int b = blockIdx.y;
  const long long nrow = rt.rows[b];
  const int S = rt.mask[b];
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nrow; r += (long long)gridDim.x * blockDim.x) {
    long long rem = r, v = 0;
    for (int a = 0; a < dt.P; ++a) {
      const long long rd = dt.m[a] - ((S >> a) & 1);
      v += (rem % rd) * dt.stride[a];
      rem /= rd;
    }
    const long long base = dt.plane + v;
    double d = 0;
    for (int f = 0; f < bt.nsub[b]; ++f) {
      const double t = theta_ghosted[base + bt.off[b][f]];
      d += (__popc(bt.sub[b][f]) & 1) ? -t : t;
    }
    rows[rt.row_off[b] + r] = bt.scale[b] * d;
  }
}

// out = D^T * rows  (rows compact, out ghosted slab of doubles)
__global__ void k_apply_Dt(const __grid_constant__ DimTab dt, const __grid_constant__ BlockTab bt,
                           const __grid_constant__ RowTab rt, const double *rows, double *out_ghosted) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zl = blockIdx.y;
  if (q >= dt.plane) return;
  int lo_ok, hi_ok;
  boundary_masks(dt, q, zl, lo_ok, hi_ok);
  // coordinates again, for the compact row index
  long long idx[MVTV_MAXP];
  {
    long long rem = q;
    for (int a = 0; a < dt.P - 1; ++a) { idx[a] = rem % dt.m[a]; rem /= dt.m[a]; }
    idx[dt.P - 1] = zl;
  }
  double total = 0;
  for (int b = 0; b < bt.K; ++b) {
    const int S = bt.mask[b];
    double acc = 0;
    for (int j = 0; j < bt.nsub[b]; ++j) {
      const int e = bt.sub[b][j];
      if ((e & ~lo_ok) != 0) continue;
      if (((S & ~e) & ~hi_ok) != 0) continue;
      long long r = 0;
      for (int a = 0; a < dt.P; ++a) r += (idx[a] - ((e >> a) & 1)) * rt.rstride[b][a];
      const double w = rows[rt.row_off[b] + r];
      acc += (__popc(e) & 1) ? -w : w;
    }
    total += bt.scale[b] * acc;
  }
  out_ghosted[(long long)(zl + 1) * dt.plane + q] = total;
}

}  // namespace mvtv
