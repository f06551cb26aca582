// Run-time binding of the handful of NCCL entry points the sharded executor
// needs (point-to-point exchange of half shards over NVLink / NVSwitch).
// dlopen()ed on first use so that the single-GPU library has no link-time
// dependency on NCCL; inside a torch process the already-loaded libnccl.so.2
// (the one torch.distributed uses) is picked up.
#pragma once
#include <dlfcn.h>
#include <stddef.h>

#include <cuda_runtime.h>

namespace qdc {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // 0 == ncclSuccess
enum { kNcclChar = 0, kNcclFloat64 = 8 };
enum { kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  const char* load() {
    if (handle) return nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) return "libnccl.so.2 not found (dlopen failed)";
#define QDC_SYM(field, name)                                  \
  *(void**)(&field) = dlsym(handle, name);                     \
  if (!field) return "NCCL symbol " name " not found";
    QDC_SYM(GetUniqueId, "ncclGetUniqueId")
    QDC_SYM(CommInitRank, "ncclCommInitRank")
    QDC_SYM(CommDestroy, "ncclCommDestroy")
    QDC_SYM(Send, "ncclSend")
    QDC_SYM(Recv, "ncclRecv")
    QDC_SYM(GroupStart, "ncclGroupStart")
    QDC_SYM(GroupEnd, "ncclGroupEnd")
    QDC_SYM(AllReduce, "ncclAllReduce")
    QDC_SYM(AllGather, "ncclAllGather")
    QDC_SYM(GetErrorString, "ncclGetErrorString")
#undef QDC_SYM
    return nullptr;
  }
};

inline NcclApi& nccl() {
  static NcclApi api;
  return api;
}

}  // namespace qdc
