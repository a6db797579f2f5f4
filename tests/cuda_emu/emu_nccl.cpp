// emu_nccl.cpp -- TEST INFRASTRUCTURE ONLY: an in-process stand-in for the NCCL entry points csrc/solver.cu resolves with dlsym,
// for the multi-rank emulation (tests/cuda_emu/emu_multi_check.py): every rank is a host thread of one process, "device" memory is
// host memory, a communicator is a rendezvous object found by its unique id.  Collectives are blocking and barrier based;
// send / recv copy through per-(source, destination) mailboxes, so grouped exchanges cannot deadlock.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

#include "fake/nccl.h"

namespace {
struct Shared {
  int world = 0, joined = 0, left = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long gen = 0;
  std::vector<const void *> send;
  std::map<std::pair<int, int>, std::deque<std::vector<unsigned char>>> mail;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long long g = gen;
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
std::mutex g_mu;
std::map<unsigned long long, Shared *> g_comms;
unsigned long long g_next_id = 1;
size_t tsize(ncclDataType_t t) { return t == ncclFloat64 ? 8 : (t == ncclFloat32 ? 4 : 1); }
struct Op { bool send; const void *src; void *dst; size_t bytes; int peer; };
thread_local int t_group = 0;
thread_local std::vector<std::pair<struct ncclComm *, Op>> t_ops;
}  // namespace

struct ncclComm { Shared *sh; int rank; };

static void do_send(ncclComm *c, const Op &o) {
  std::lock_guard<std::mutex> lk(c->sh->mu);
  c->sh->mail[{c->rank, o.peer}].emplace_back((const unsigned char *)o.src, (const unsigned char *)o.src + o.bytes);
  c->sh->cv.notify_all();
}
static void do_recv(ncclComm *c, const Op &o) {
  std::unique_lock<std::mutex> lk(c->sh->mu);
  auto &q = c->sh->mail[{o.peer, c->rank}];
  c->sh->cv.wait(lk, [&] { return !q.empty(); });
  std::memcpy(o.dst, q.front().data(), o.bytes);
  q.pop_front();
}
static void flush_ops() {
  for (auto &e : t_ops) if (e.second.send) do_send(e.first, e.second);
  for (auto &e : t_ops) if (!e.second.send) do_recv(e.first, e.second);
  t_ops.clear();
}

extern "C" {
const char *ncclGetErrorString(ncclResult_t) { return "emu_nccl error"; }
ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::memset(id, 0, sizeof(*id));
  const unsigned long long v = g_next_id++;
  std::memcpy(id->internal, &v, sizeof(v));
  return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t *out, int world, ncclUniqueId id, int rank) {
  unsigned long long key;
  std::memcpy(&key, id.internal, sizeof(key));
  Shared *sh;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    Shared *&slot = g_comms[key];
    if (!slot) { slot = new Shared(); slot->world = world; slot->send.assign((size_t)world, nullptr); }
    sh = slot;
  }
  *out = new ncclComm{sh, rank};
  sh->barrier();   // like NCCL: returns when every rank has joined
  return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }   // the rendezvous object is leaked on purpose (tests)
ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c, void *) {
  Shared *sh = c->sh;
  sh->send[(size_t)c->rank] = send;
  sh->barrier();
  std::vector<unsigned char> tmp(count * tsize(t));
  for (size_t i = 0; i < count; ++i) {   // rank order: the same bits on every rank
    if (t == ncclFloat64) {
      double a = ((const double *)sh->send[0])[i];
      for (int j = 1; j < sh->world; ++j) { const double b = ((const double *)sh->send[(size_t)j])[i]; a = (op == ncclMax) ? (b > a ? b : a) : a + b; }
      ((double *)tmp.data())[i] = a;
    } else if (t == ncclFloat32) {
      float a = ((const float *)sh->send[0])[i];
      for (int j = 1; j < sh->world; ++j) { const float b = ((const float *)sh->send[(size_t)j])[i]; a = (op == ncclMax) ? (b > a ? b : a) : a + b; }
      ((float *)tmp.data())[i] = a;
    } else {
      return ncclUnhandledCudaError;
    }
  }
  sh->barrier();   // everybody has read every send buffer (in-place calls overwrite them next)
  std::memcpy(recv, tmp.data(), tmp.size());
  return ncclSuccess;
}
ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t c, void *) {
  Shared *sh = c->sh;
  const size_t bytes = count * tsize(t);
  sh->send[(size_t)c->rank] = send;
  sh->barrier();
  std::vector<unsigned char> tmp(bytes * (size_t)sh->world);
  for (int j = 0; j < sh->world; ++j) std::memcpy(tmp.data() + bytes * (size_t)j, sh->send[(size_t)j], bytes);
  sh->barrier();
  std::memcpy(recv, tmp.data(), tmp.size());
  return ncclSuccess;
}
ncclResult_t ncclGroupStart() { ++t_group; return ncclSuccess; }
ncclResult_t ncclGroupEnd() { if (--t_group == 0) flush_ops(); return ncclSuccess; }
ncclResult_t ncclSend(const void *src, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void *) {
  t_ops.push_back({c, Op{true, src, nullptr, count * tsize(t), peer}});
  if (t_group == 0) flush_ops();
  return ncclSuccess;
}
ncclResult_t ncclRecv(void *dst, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void *) {
  t_ops.push_back({c, Op{false, nullptr, dst, count * tsize(t), peer}});
  if (t_group == 0) flush_ops();
  return ncclSuccess;
}
}
