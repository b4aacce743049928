// nccl_shim.h -- NCCL bound at run time with dlopen (libnccl.so.2; the copy torch already loaded is reused when the
// caller is a torch.distributed process), so libaha_b200.so has no build-time NCCL dependency.  Used only for the
// tensor-parallel exchange steps: all-reduce(sum) of the o_proj and down_proj partial outputs -- the
// reference has no distributed code at all.
#pragma once
#include <dlfcn.h>

#include <cstring>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

namespace aha {

struct NcclApi {
    typedef struct ncclComm* comm_t;
    struct unique_id { char internal[128]; };
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t /*sendcount*/, int /*dtype*/, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t /*count*/, int /*dtype*/, int /*root*/, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* handle = nullptr;
    static constexpr int kFloat32 = 7, kInt8 = 0, kSum = 0;

    static NcclApi& get() {
        static NcclApi api;
        if (!api.handle) {
            const char* names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char* n : names) {
                api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (api.handle) break;
            }
            if (!api.handle) throw std::runtime_error("tensor parallelism needs NCCL: dlopen(libnccl.so.2) failed");
            auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + s); return p; };
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
            api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
            api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        }
        return api;
    }
    void check(int rc, const char* what) const {
        if (rc != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(rc) : "?"));
    }
};

}  // namespace aha
