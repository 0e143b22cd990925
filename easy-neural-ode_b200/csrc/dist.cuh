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

struct DistCtx {
  int world = 1, rank = 0;
  NcclComm comm = nullptr;
  NcclApi api;
  bool active() const { return world > 1 && comm != nullptr; }

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
