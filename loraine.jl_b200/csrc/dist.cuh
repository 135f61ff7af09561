// NCCL plumbing (one process per GPU).  libnccl is resolved at run time with dlopen so that the library has no
// link-time dependency on a particular NCCL build (torch bundles 2.28.9, the system has 2.27.3).
#pragma once
#include "common.cuh"
#include <nccl.h>

namespace lrn {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
const NcclApi& nccl_api();   // throws std::runtime_error when libnccl cannot be loaded

struct DistCtx {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

struct CholWork;
// Distributed right-looking Cholesky, 1-D block-cyclic by column panels of width pw (panel p -> rank p % world).
// On entry every rank holds its OWN panels of the lower triangle of A (other panels: don't care); on return every rank
// holds the complete factor L and all inverse diagonal blocks (panels are broadcast with ncclBroadcast as they are
// finished, the receivers store them), so triangular solves run replicated without further communication.
void cholesky_dist(double* A, int n, int lda, CholWork& work, DistCtx& ctx, int pw, DevBuf<double>& panelbuf, cudaStream_t st);
// in-place sum of a device buffer over all ranks
void dist_allreduce_sum(double* buf, size_t count, DistCtx& ctx, cudaStream_t st);

#define LRN_NCCL(call)                                                                         \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess) {                                                              \
            char buf__[512];                                                                   \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,        \
                     lrn::nccl_api().GetErrorString ? lrn::nccl_api().GetErrorString(r__) : "nccl error"); \
            throw std::runtime_error(buf__);                                                   \
        }                                                                                      \
    } while (0)

}  // namespace lrn
