// Multi-GPU plumbing: one process per GPU, NCCL over NVLink.  The NCCL entry points are resolved at run
// time from the libnccl.so.2 that PyTorch has already loaded (dlopen NOLOAD / explicit path), so
// libeno.so has no link-time dependency on NCCL and single-GPU use never touches it.
//
// What travels per solver iteration:
//   forward : all-gather of the per-sample fp64 partial sums of (err/tol)^2  (kC * B_global doubles)
//   adjoint : ONE all-gather of a packed per-rank record (the same partial sums + the per-sample t_bar pieces
//             of all six stages, adj_host.cuh) and ONE all-reduce of the four stage combinations of the
//             params_bar block (4 P floats): everything downstream of the batch sum is linear.
// After the exchange every rank holds bit-identical inputs and runs the same one-CTA controller, so the
// global accept / reject sequence is the single-device sequence for the concatenated batch.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <string>

namespace eno {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclSum = 0, kNcclFloat32 = 7, kNcclFloat64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  bool load(const char* path, std::string* err) {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib && path && *path) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      if (err) *err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return false;
    }
#define ENO_SYM(field, name)                                                  \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                \
  if (!field) { if (err) *err = std::string("libnccl lacks ") + name; return false; }
    ENO_SYM(GetUniqueId, "ncclGetUniqueId")
    ENO_SYM(CommInitRank, "ncclCommInitRank")
    ENO_SYM(CommDestroy, "ncclCommDestroy")
    ENO_SYM(AllReduce, "ncclAllReduce")
    ENO_SYM(AllGather, "ncclAllGather")
    ENO_SYM(GroupStart, "ncclGroupStart")
    ENO_SYM(GroupEnd, "ncclGroupEnd")
    ENO_SYM(GetErrorString, "ncclGetErrorString")
#undef ENO_SYM
    return true;
  }
};

// ---- exchange over NVLink peer memory (no NCCL call per solver iteration) -----------------------------------------
// Every rank owns one "symmetric" buffer (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by every peer), laid out
// identically on all ranks:
//   [ flags: 2 kinds x 8 ranks, u64 | seq (local) | timeout word | part[2 parities][8 ranks][2] doubles | PeerView | data ... ]
// A rank SIGNALS by storing a monotonically increasing sequence number into slot [kind][its rank] of EVERY peer's buffer
// (st.release.sys after a system fence) and WAITS by spinning on its own, local, slots - so polling never crosses
// NVLink.  Small payloads (per-sample norm partials, the 24 KB adjoint record, the per-rank norm scalar) are PUSHED
// into the peers' buffers before the signal; the large one (the params_bar stage combinations, 4 P floats) stays
// where it was produced and every rank PULLS only the 1/world slice it owns (ld.global.cv), in rank order, so sums are
// deterministic.  Payload areas are double-buffered by the parity of the sequence number: a rank can be at most one
// exchange ahead of the slowest peer.
constexpr int kPeerMax = 8;
constexpr size_t kPeerOffFlags = 0;            // u64 [2][8]
constexpr size_t kPeerOffSeq = 128;            // u64, local use
constexpr size_t kPeerOffTimeout = 136;        // int, local use: a wait gave up (peer lost)
constexpr size_t kPeerOffPart = 256;           // double [2][8][2]
constexpr size_t kPeerOffView = 512;           // this rank's PeerView (device-resident copy read by the kernels)
constexpr size_t kPeerOffFwd = 1024;           // forward: double [2 parities][2 arrays][n_rp]
constexpr size_t kPeerOffAdj = (size_t)1 << 20;  // adjoint: xrec [2][world][6 n_loc] doubles, then pcomb [2][4 P] floats

struct PeerView {
  int world = 1, rank = 0;
  unsigned char* base[kPeerMax] = {nullptr};   // base[r]: rank r's symmetric buffer as mapped in THIS process
  size_t bytes = 0;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// threads 0..world-1 of the calling CTA; the CTA's earlier writes (any thread) must be ordered before by __syncthreads
__device__ __forceinline__ void peer_signal(const PeerView& pv, int kind, unsigned long long seq) {
  if ((int)threadIdx.x < pv.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(pv.base[threadIdx.x] + kPeerOffFlags) + kind * kPeerMax + pv.rank, seq);
  }
}
// every thread of the CTA returns once all ranks have signalled `seq` (or a ~10 s timeout has been recorded)
__device__ __forceinline__ void peer_wait(const PeerView& pv, int kind, unsigned long long seq) {
  if ((int)threadIdx.x < pv.world) {
    const unsigned long long* f =
        reinterpret_cast<const unsigned long long*>(pv.base[pv.rank] + kPeerOffFlags) + kind * kPeerMax + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < seq) {
      if (clock64() - t0 > 20000000000LL) {   // a peer is gone: record it and stop waiting rather than hang the GPU
        *reinterpret_cast<volatile int*>(pv.base[pv.rank] + kPeerOffTimeout) = 1;
        break;
      }
    }
  }
  __syncthreads();
}
#endif

struct DistCtx {
  int world = 1, rank = 0;
  NcclComm comm = nullptr;
  NcclApi api;
  PeerView peer;            // mapped when eno_peer_attach has run; peer.base[rank] == nullptr otherwise
  void* peer_mine = nullptr;
  // Uneven shards (eno_comm_set_shards): every rank's records are laid out for `slot_rows` rows (the largest shard); a
  // rank with fewer rows leaves the tail of its records zero.  rows_global = rows of the concatenated batch.
  int slot_rows = 0;
  long long rows_global = 0;
  int slot(int B) const { return (active() && slot_rows > 0) ? slot_rows : B; }
  long long rows(int B) const { return (active() && rows_global > 0) ? rows_global : (long long)B * (active() ? world : 1); }
  bool active() const { return world > 1 && comm != nullptr; }
  bool peer_ready() const { return active() && peer.base[rank] != nullptr; }

  // in-place all-gather of equal per-rank chunks: buf holds world * count elements, this rank's at rank * count
  template <typename T>
  int gather(T* buf, size_t count, cudaStream_t st) const {
    const int type = sizeof(T) == 8 ? kNcclFloat64 : kNcclFloat32;
    return api.AllGather(buf + (size_t)rank * count, buf, count, type, comm, st);
  }
  int allreduce(float* buf, size_t count, cudaStream_t st) const {
    return api.AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, comm, st);
  }
};

}  // namespace eno
