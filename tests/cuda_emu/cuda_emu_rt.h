// cuda_emu_rt.h -- TEST INFRASTRUCTURE ONLY: host-side stand-ins for the part of the CUDA runtime API that csrc/solver.cu and
// csrc/setup.cu call, so that the library's host code (after tests/cuda_emu/emu_translate.py has rewritten its <<<>>> launches)
// compiles with g++ against the SIMT emulator of cuda_emu.h.  "Device" memory is host memory, streams are synchronous,
// events are wall-clock stamps, one device with EMU_NSM (default 4) multiprocessors; IPC handles carry raw pointers (ranks = threads of one process, emu_nccl.cpp).
#pragma once
#include <chrono>

#include "cuda_emu.h"

struct cuda_emu_event { std::chrono::steady_clock::time_point t; };
typedef cuda_emu_event *cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocPortable = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaErrorNotSupported = 801 };

inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) {
  const char *e = std::getenv("EMU_NSM");
  *v = e ? std::atoi(e) : 4;
  return cudaSuccess;
}
template <typename PP>
inline cudaError_t cudaMalloc(PP **p, size_t bytes) {
  *p = (PP *)std::aligned_alloc(256, (bytes + 255) / 256 * 256 + 256);
  if (*p) std::memset((void *)*p, 0xA5, bytes);   // "uninitialised" device memory is visibly garbage
  return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
template <typename PP>
inline cudaError_t cudaMallocHost(PP **p, size_t bytes) { *p = (PP *)std::aligned_alloc(256, (bytes + 255) / 256 * 256 + 256); return *p ? cudaSuccess : 2; }
template <typename PP>
inline cudaError_t cudaHostAlloc(PP **p, size_t bytes, unsigned) { return cudaMallocHost(p, bytes); }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuda_emu_event(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
// occupancy: EMU_OCC resident CTAs per multiprocessor (default 2) whatever the kernel
template <typename F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *occ, F, int, size_t) {
  const char *e = std::getenv("EMU_OCC");
  *occ = e ? std::atoi(e) : 2;
  return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// ranks of the multi-rank emulation are threads of one process: an IPC handle carries the raw pointer
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof(*p)); return *p ? cudaSuccess : cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
