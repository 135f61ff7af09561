#include "dist.cuh"
#include "solver.cuh"
#include <dlfcn.h>

namespace lrn {

const NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* hnd = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!hnd) hnd = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (hnd) {
            auto sym = [&](const char* n) { return dlsym(hnd, n); };
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
            api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
            api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.Broadcast && api.AllReduce && api.AllGather;
        }
    }
    if (!api.ok) throw std::runtime_error("libnccl.so.2 could not be loaded (needed for multi-GPU runs)");
    return api;
}

}  // namespace lrn

using namespace lrn;

extern "C" {

int32_t lrn_dist_unique_id(void* out128) {
    if (!out128) return LRN_ERR_ARG;
    try {
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
        ncclUniqueId id;
        if (nccl_api().GetUniqueId(&id) != ncclSuccess) return LRN_ERR_NCCL;
        std::memcpy(out128, &id, sizeof id);
        return LRN_OK;
    } catch (...) {
        return LRN_ERR_NCCL;
    }
}

int32_t lrn_dist_init(lrn_handle_t h, int32_t rank, int32_t world, const void* unique_id128) {
    if (!h || !unique_id128 || world < 1 || rank < 0 || rank >= world) return LRN_ERR_ARG;
    try {
        LRN_CUDA(cudaSetDevice(h->device));
        ncclUniqueId id;
        std::memcpy(&id, unique_id128, sizeof id);
        auto* ctx = new DistCtx();
        ctx->rank = rank;
        ctx->world = world;
        LRN_NCCL(nccl_api().CommInitRank(&ctx->comm, world, id, rank));
        h->nccl = ctx;
        h->rank = rank;
        h->world = world;
        return LRN_OK;
    } catch (const std::exception& e) {
        h->err = e.what();
        return LRN_ERR_NCCL;
    }
}

}  // extern "C"
