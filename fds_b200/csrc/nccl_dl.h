// NCCL bound at run time (dlopen): the library has no link-time dependency on NCCL, a single-GPU
// client never loads it, and under PyTorch the process's already-loaded libnccl.so.2 is the one used.
// Only the handful of entry points of the N-sharded segment loop are bound: ring halo send/recv, the
// (m+2)^2 Gram all-reduce and the all-gather of the per-rank objective partials
// (reference: the MPI exchanges of tests/solvers/mock_fun3d/fun3d:29-37 and
// pascal_lite/operators/plinalg.py:19-64).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdlib>
#include <string>

namespace fds {

struct NcclUniqueId { char internal[128]; };      // NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;
constexpr int kNcclFloat64 = 8;                   // ncclFloat64
constexpr int kNcclSum = 0;                       // ncclSum

struct NcclApi {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*CommCount)(NcclComm, int*) = nullptr;
    int (*CommUserRank)(NcclComm, int*) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

// the process-wide binding (loaded on first use); nullptr + message when NCCL is not there
inline const NcclApi* nccl_api(std::string* err) {
    static NcclApi api;
    static bool tried = false;
    static std::string why;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("FDS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !nm[0]) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
            why = dlerror();
        }
        if (api.handle) {
            bool ok = true;
#define FDS_NCCL_SYM(field, name)                                                  \
    do {                                                                           \
        *reinterpret_cast<void**>(&api.field) = dlsym(api.handle, name);           \
        if (!api.field) { ok = false; why = std::string("missing symbol ") + name; } \
    } while (0)
            FDS_NCCL_SYM(GetVersion, "ncclGetVersion");
            FDS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
            FDS_NCCL_SYM(CommInitRank, "ncclCommInitRank");
            FDS_NCCL_SYM(CommDestroy, "ncclCommDestroy");
            FDS_NCCL_SYM(CommCount, "ncclCommCount");
            FDS_NCCL_SYM(CommUserRank, "ncclCommUserRank");
            FDS_NCCL_SYM(Send, "ncclSend");
            FDS_NCCL_SYM(Recv, "ncclRecv");
            FDS_NCCL_SYM(AllReduce, "ncclAllReduce");
            FDS_NCCL_SYM(AllGather, "ncclAllGather");
            FDS_NCCL_SYM(GroupStart, "ncclGroupStart");
            FDS_NCCL_SYM(GroupEnd, "ncclGroupEnd");
            FDS_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef FDS_NCCL_SYM
            if (!ok) { dlclose(api.handle); api.handle = nullptr; }
        }
    }
    if (!api.handle) {
        if (err) *err = "NCCL not available (" + why + "); set FDS_NCCL_LIB to libnccl.so.2";
        return nullptr;
    }
    return &api;
}

}  // namespace fds
