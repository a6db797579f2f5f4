// cuda_emu.h -- a small SIMT emulator, TEST INFRASTRUCTURE ONLY: runs the library's CUDA kernel SOURCE (csrc/*.cuh) on the
// CPU so that kernel logic can be checked without a GPU (index arithmetic, halo handling, warp-shuffle exchanges, barriers,
// the deterministic grid reduction).  It says nothing about performance, memory-model races or PTX-level behaviour.
//
// Every CUDA thread of a block is a ucontext fiber; blocks run one after another.  __syncthreads() and the implicit
// convergence of __shfl_*_sync are barriers between fibers (threads that have exited count as arrived); a barrier that
// can never complete (divergent shuffle, missing thread) aborts with a message instead of hanging.  __shared__ variables
// become function-local statics (blocks are sequential, so a static is private to the running block); dynamic shared
// memory is one buffer handed out by MVTV_DYN_SMEM.  cp.async and the .sys loads / stores have plain-C++ bodies under
// MVTV_CUDA_EMU in kernels.cuh.
#pragma once
#ifndef MVTV_CUDA_EMU
#define MVTV_CUDA_EMU 1
#endif
#ifndef __CUDACC__
#define __CUDACC__ 1
#endif
#include <setjmp.h>
#include <time.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local   // one emulated device per host thread (multi-rank emulation: one thread per rank)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
struct float2 { float x, y; };
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
inline const char *cudaGetErrorString(cudaError_t) { return "cuda_emu"; }

namespace cuda_emu {

struct Fiber {
  ucontext_t ctx;          // only for the first entry; later switches are _setjmp / _longjmp (no signal-mask system call)
  jmp_buf jb;
  bool started = false;
  char *stack = nullptr;   // from the pool below: allocated once, reused by every block (no re-zeroing)
  bool done = false;
  uint3 tid{0, 0, 0};
  int linear = 0;
};
struct Barrier {
  int arrived = 0;
  unsigned long long gen = 0;
};
struct State {
  dim3 grid, block;
  uint3 bid{0, 0, 0};
  Fiber *cur = nullptr;
  ucontext_t sched;
  jmp_buf sched_jb;
  std::vector<Fiber> fibers;
  int alive = 0;
  Barrier blockbar;
  std::vector<Barrier> warpbar;
  std::vector<int> warp_alive;
  std::vector<unsigned long long> warpbuf;   // [warp][32] 8-byte slots for shuffles
  std::vector<unsigned char> smem;
  std::function<void()> body;
  unsigned long long progress = 0;
  std::vector<char *> stack_pool;
};
static const size_t kStackBytes = 256 * 1024;
inline State &g() {
  static thread_local State s;   // per host thread: several ranks can run their kernels concurrently
  return s;
}
inline unsigned char *dyn_smem() { return g().smem.data(); }

inline void yield() {
  State &s = g();
  if (_setjmp(s.cur->jb) == 0) _longjmp(s.sched_jb, 1);
}
// generic barrier over `count()` participants; count is re-evaluated while waiting because threads may exit meanwhile
template <typename CountFn>
inline void barrier_wait(Barrier &b, CountFn count) {
  State &s = g();
  const unsigned long long gen = b.gen;
  b.arrived += 1;
  for (;;) {
    if (b.gen != gen) return;
    if (b.arrived >= count()) {
      b.arrived = 0;
      b.gen += 1;
      s.progress += 1;
      return;
    }
    yield();
  }
}
inline void sync_block() {
  State &s = g();
  barrier_wait(s.blockbar, [&] { return s.alive; });
}
inline void sync_warp() {
  State &s = g();
  const int w = s.cur->linear >> 5;
  barrier_wait(s.warpbar[(size_t)w], [&, w] { return s.warp_alive[(size_t)w]; });
}
template <typename T>
inline T shfl_from(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  State &s = g();
  const int lane = s.cur->linear & 31, w = s.cur->linear >> 5;
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  s.warpbuf[(size_t)w * 32 + lane] = bits;
  sync_warp();
  T r = v;
  if (src >= 0 && src < 32) {
    const unsigned long long b2 = s.warpbuf[(size_t)w * 32 + src];
    std::memcpy(&r, &b2, sizeof(T));
  }
  sync_warp();   // nobody overwrites a slot before everybody has read
  return r;
}
inline void fiber_entry() {
  State &s = g();
  s.body();
  Fiber *f = s.cur;
  f->done = true;
  s.alive -= 1;
  s.warp_alive[(size_t)(f->linear >> 5)] -= 1;
  s.progress += 1;
  _longjmp(s.sched_jb, 1);
}

// launch<<<grid, block, smem>>>: `body` calls the kernel with its arguments
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body) {
  State &s = g();
  s.grid = grid;
  s.block = block;
  s.body = body;
  const int nt = (int)(block.x * block.y * block.z);
  if (nt < 1 || nt > 1024) { std::fprintf(stderr, "cuda_emu: block size %d\n", nt); std::abort(); }
  const int nwarps = (nt + 31) / 32;   // the last warp may be partial (e.g. the <<<1, 1>>> commit kernels)
  s.smem.assign(smem_bytes + 64, 0);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bid = uint3{bx, by, bz};
        s.fibers.assign((size_t)nt, Fiber());
        s.alive = nt;
        s.blockbar = Barrier();
        s.warpbar.assign((size_t)nwarps, Barrier());
        s.warp_alive.assign((size_t)nwarps, 32);
        s.warp_alive[(size_t)nwarps - 1] = nt - 32 * (nwarps - 1);
        s.warpbuf.assign((size_t)nwarps * 32, 0);
        for (int t = 0; t < nt; ++t) {
          Fiber &f = s.fibers[(size_t)t];
          f.linear = t;
          f.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
          while (s.stack_pool.size() <= (size_t)t) s.stack_pool.push_back((char *)std::malloc(kStackBytes));
          f.stack = s.stack_pool[(size_t)t];
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = &s.sched;
          makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        int idle_rounds = 0;
        while (s.alive > 0) {
          const unsigned long long before = s.progress;
          for (int t = 0; t < nt; ++t) {
            Fiber &f = s.fibers[(size_t)t];
            if (f.done) continue;
            s.cur = &f;
            if (_setjmp(s.sched_jb) == 0) {
              if (!f.started) {
                f.started = true;
                setcontext(&f.ctx);
              } else {
                _longjmp(f.jb, 1);
              }
            }
          }
          idle_rounds = (s.progress == before) ? idle_rounds + 1 : 0;
          if (idle_rounds > 4) {
            std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %d threads wait at a barrier nobody else reaches\n", bx, by, bz, s.alive);
            std::abort();
          }
        }
      }
}

}  // namespace cuda_emu

#define threadIdx (::cuda_emu::g().cur->tid)
#define blockIdx (::cuda_emu::g().bid)
#define blockDim (::cuda_emu::g().block)
#define gridDim (::cuda_emu::g().grid)
#define MVTV_DYN_SMEM(name) unsigned char *name = ::cuda_emu::dyn_smem()

inline void __syncthreads() { ::cuda_emu::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ::cuda_emu::sync_warp(); }
inline void __threadfence() {}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }   // ranks are host threads: a real fence
inline long long clock64() {   // ~2 ticks per nanosecond, so that peer_spin's time-out (kernels.cuh) also ends a stuck emulated wait
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 2ll * ((long long)ts.tv_sec * 1000000000ll + ts.tv_nsec);
}
template <typename T> inline T __ldcg(const T *p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned atomicInc(unsigned *addr, unsigned val) {
  const unsigned old = *addr;
  *addr = (old >= val) ? 0u : old + 1u;
  return old;
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d) { const int lane = ::cuda_emu::g().cur->linear & 31; return ::cuda_emu::shfl_from(v, lane - d >= 0 ? lane - d : -1); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) { const int lane = ::cuda_emu::g().cur->linear & 31; return ::cuda_emu::shfl_from(v, lane + d < 32 ? lane + d : -1); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { const int lane = ::cuda_emu::g().cur->linear & 31; return ::cuda_emu::shfl_from(v, lane ^ m); }
using std::max;
using std::min;
