// Minimal run-time binding of the NCCL entry points this library uses (the torch wheel ships
// libnccl.so.2; there is no link-time dependency so the single-GPU path works without it).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <stdlib.h>

namespace recs {

struct NcclApi {
    typedef struct ncclComm* comm_t;
    struct unique_id {
        char internal[128];
    };
    // ncclDataType_t: ncclFloat32 == 7; ncclResult_t: ncclSuccess == 0.
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* handle = nullptr;

    bool load(const char** why) {
        if (handle) return true;
        const char* candidates[4] = {getenv("RECS_GPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; i < 3 && !handle; ++i) {
            if (!candidates[i] || !candidates[i][0]) continue;
            // prefer a copy already mapped by the process (torch loads its bundled one)
            handle = dlopen(candidates[i], RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (!handle) handle = dlopen(candidates[i], RTLD_NOW | RTLD_GLOBAL);
        }
        if (!handle) {
            *why = "libnccl.so.2 not found (set RECS_GPU_NCCL_LIB or import torch first)";
            return false;
        }
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(handle, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(handle, "ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(handle, "ncclAllGather"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(handle, "ncclBroadcast"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(handle, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(handle, "ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString || !Broadcast || !GroupStart || !GroupEnd) {
            *why = "libnccl is missing a required symbol";
            return false;
        }
        return true;
    }
};

}  // namespace recs
