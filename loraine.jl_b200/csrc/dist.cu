#include "dist.cuh"
#include "solver.cuh"
#include "chol.cuh"
#include "gemm.cuh"
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

namespace {
__global__ void k_pack_panel(const double* __restrict__ A, int lda, int rows, int w, double* __restrict__ P, int ldp, int unpack) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= rows || j >= w) return;
    if (unpack) const_cast<double*>(A)[(size_t)j * lda + i] = P[(size_t)j * ldp + i];
    else P[(size_t)j * ldp + i] = A[(size_t)j * lda + i];
}
__global__ void k_max_int(int* a, const int* b) { if (*b != 0 && (*a == 0 || *b < *a)) *a = *b; }
}  // namespace

void dist_allreduce_sum(double* buf, size_t count, DistCtx& ctx, cudaStream_t st) {
    LRN_NCCL(nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, ctx.comm, st));
}

void cholesky_dist(double* A, int n, int lda, CholWork& work, DistCtx& ctx, int pw, DevBuf<double>& panelbuf, cudaStream_t st) {
    work.ensure(n);
    LRN_REQUIRE(pw % CHOL_DB == 0, "panel width must be a multiple of 64");
    const int npan = (int)cdiv(n, pw), ldp = pad_ld(n);
    const size_t need = (size_t)ldp * pw + (size_t)(pw / CHOL_DB) * CHOL_DB * CHOL_DB + 8;
    if (panelbuf.n < need) panelbuf.alloc(need);
    int* info = work.info_ptr();
    LRN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    for (int p = 0; p < npan; p++) {
        const int c0 = p * pw, w = (n - c0 < pw) ? (n - c0) : pw, rows = n - c0, owner = p % ctx.world;
        double* Ap = A + (size_t)c0 * lda + c0;
        double* dk = work.dinv.p + (size_t)(c0 / CHOL_DB) * CHOL_DB * CHOL_DB;
        const int nd = (int)cdiv(w, CHOL_DB) * CHOL_DB * CHOL_DB;
        double* dbuf = panelbuf.p + (size_t)ldp * pw;
        dim3 grid((unsigned)cdiv(rows, 256), (unsigned)w);
        if (ctx.rank == owner) {
            cholesky_panel(Ap, rows, w, lda, dk, info, c0, st);
            k_pack_panel<<<grid, 256, 0, st>>>(Ap, lda, rows, w, panelbuf.p, ldp, 0);
            LRN_CHECK_LAUNCH();
            LRN_CUDA(cudaMemcpyAsync(dbuf, dk, (size_t)nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        // one broadcast carries the panel and the inverse diagonal blocks (contiguous in panelbuf)
        LRN_NCCL(nccl_api().Broadcast(panelbuf.p, panelbuf.p, (size_t)ldp * pw + nd, ncclDouble, owner, ctx.comm, st));
        if (ctx.rank != owner) {
            k_pack_panel<<<grid, 256, 0, st>>>(Ap, lda, rows, w, panelbuf.p, ldp, 1);
            LRN_CHECK_LAUNCH();
            LRN_CUDA(cudaMemcpyAsync(dk, dbuf, (size_t)nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        for (int q = p + 1; q < npan; q++) {
            if (q % ctx.world != ctx.rank) continue;
            const int q0 = q * pw, wq = (n - q0 < pw) ? (n - q0) : pw;
            const double* Pq = panelbuf.p + (q0 - c0);
            gemm_nt(st, n - q0, wq, w, -1.0, Pq, ldp, Pq, ldp, 1.0, A + (size_t)q0 * lda + q0, lda);
        }
    }
    // the first failing pivot index is known to the owner of that panel only: take the smallest non-zero over ranks
    DevBuf<int>& tmp = work.info;   // scratch int lives next to the flag when an external flag is used
    int* all = nullptr;
    LRN_CUDA(cudaMalloc(&all, sizeof(int) * ctx.world));
    LRN_NCCL(nccl_api().AllGather(info, all, 1, ncclInt32, ctx.comm, st));
    for (int r = 0; r < ctx.world; r++) k_max_int<<<1, 1, 0, st>>>(info, all + r);
    LRN_CUDA(cudaStreamSynchronize(st));
    cudaFree(all);
    (void)tmp;
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
        h->dist_pw = (h->n_var >= 16384) ? 512 : (h->n_var >= 4096 ? 256 : 128);
        return LRN_OK;
    } catch (const std::exception& e) {
        h->err = e.what();
        return LRN_ERR_NCCL;
    }
}

}  // extern "C"
