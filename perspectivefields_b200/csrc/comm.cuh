// Multi-GPU result gather over NCCL point-to-point: one process per GPU; the images of an inference_batch list
// are sharded over the ranks and the per-image result tensors are gathered to one rank with grouped ncclSend / ncclRecv on the
// caller's stream.  NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy torch already
// loaded), so libpf_b200.so has no link-time dependency on it; the handful of declarations below is NCCL's stable C ABI.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <mutex>

namespace pf {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { kNcclSuccess = 0, kNcclUint8 = 1 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  const char* error = nullptr;   // why loading failed
};

inline const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.error = "libnccl.so.2 could not be loaded (import torch first, or put NCCL on the library path)"; return; }
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p && !api.error) api.error = "libnccl.so.2 lacks a required symbol"; return p; };
    api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  });
  return api;
}

}  // namespace pf

struct pf_comm {
  int device = 0, rank = 0, nranks = 1;
  pf::NcclComm comm = nullptr;
};
