// nccl_dyn.h -- the few NCCL entry points the multi-GPU calls use, resolved at
// run time (dlopen of libnccl.so.2) so that libtrimal_cuda.so keeps linking
// nothing but the CUDA runtime: single-GPU users never need NCCL, and in a
// process that already loaded NCCL (e.g. through torch.distributed) the same
// copy is reused (same soname).  The declarations below restate the stable
// public C ABI of nccl.h (NCCL 2.x).
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

#include <mutex>

namespace tcu {

struct NcclUniqueId {
    char internal[128];
};
typedef struct ncclComm *NcclComm;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_FLOAT32 = 7 };
enum { NCCL_SUM = 0 };

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    bool ok = false;
    const char *why = "";
};

inline const NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            api.why = "libnccl.so.2 not found";
            return;
        }
        bool all = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(h, name);
            if (!p) all = false;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.ok = all;
        if (!all) api.why = "libnccl.so.2 lacks a required symbol";
    });
    return api;
}

}  // namespace tcu
