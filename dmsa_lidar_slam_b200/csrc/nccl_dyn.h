// nccl_dyn.h — NCCL bound at run time (dlopen), so that libdmsa_b200.so has no link-time dependency on it: single-GPU
// users never load NCCL, and inside a PyTorch process the already loaded libnccl.so.2 (torch's bundled copy) is reused
// instead of a second NCCL.  Only the handful of entry points the row-sharded iteration needs (one all-reduce
// of [H | g | e0^T e0] and one of the 9 line-search costs per iteration, intra-node NVLink / NVSwitch).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>
#include <cstring>
#include <string>

namespace dmsa {

// ABI-stable parts of nccl.h (NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
enum { ncclInt32 = 2, ncclFloat64 = 8 };  // ncclDataType_t
enum { ncclSum = 0, ncclMax = 2 };        // ncclRedOp_t

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;

    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return false;
        }
        auto sym = [&](const char* n) { return dlsym(lib, n); };
        GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) {
            err = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllReduce";
            lib = nullptr;
            return false;
        }
        return true;
    }
    std::string describe(ncclResult_t r) const { return GetErrorString ? std::string(GetErrorString(r)) : std::string("NCCL error ") + std::to_string((int)r); }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace dmsa
