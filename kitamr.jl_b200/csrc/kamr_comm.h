// kamr_comm.h — NCCL transport for the ghost-layer halo, loaded with dlopen so that libkamr.so has no
// link-time dependency on a particular libnccl (the host passes KAMR_NCCL_LIB or the loader finds
// libnccl.so.2).  Replaces MPI.Isend/Irecv of Parallel/Ghost.jl:203-284 with ncclSend/ncclRecv groups.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdlib>
#include <string>

namespace kamr {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int ncclFloat64 = 8;  // ncclDataType_t::ncclDouble

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    bool load(std::string& err) {
        if (handle) return true;
        const char* cands[] = {getenv("KAMR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* c : cands) {
            if (!c || !*c) continue;
            handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define KAMR_SYM(field, name)                                             \
    *(void**)(&field) = dlsym(handle, name);                              \
    if (!field) { err = std::string("missing NCCL symbol ") + name; return false; }
        KAMR_SYM(GetUniqueId, "ncclGetUniqueId");
        KAMR_SYM(CommInitRank, "ncclCommInitRank");
        KAMR_SYM(CommDestroy, "ncclCommDestroy");
        KAMR_SYM(GroupStart, "ncclGroupStart");
        KAMR_SYM(GroupEnd, "ncclGroupEnd");
        KAMR_SYM(Send, "ncclSend");
        KAMR_SYM(Recv, "ncclRecv");
        KAMR_SYM(GetErrorString, "ncclGetErrorString");
#undef KAMR_SYM
        return true;
    }
};

inline Nccl& nccl() {
    static Nccl n;
    return n;
}

}  // namespace kamr
