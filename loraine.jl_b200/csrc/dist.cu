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
    work.tinv_for = nullptr;   // the triangular solves rebuild their diagonal-block inverses from the new factor
    work.ensure(n);
    LRN_REQUIRE(pw % CHOL_DB == 0, "panel width must be a multiple of 64");
    const int npan = (int)cdiv(n, pw), ldp = pad_ld(n);
    const size_t ndmax = (size_t)(pw / CHOL_DB) * CHOL_DB * CHOL_DB;
    const size_t bufsz = (size_t)ldp * pw + ndmax + 8;
    if (panelbuf.n < 2 * bufsz) panelbuf.alloc(2 * bufsz);
    ensure_aux(work);
    cudaStream_t sp = work.aux;                     // panel stream: factor panel p+1 and broadcast it while `st` still
    cudaEvent_t evStart = work.ev[0], evU = work.ev[1];   // applies the trailing updates of panel p (one-step look-ahead)
    cudaEvent_t evB[2] = {work.ev[2], work.ev[3]}, evE[2] = {work.ev[4], work.ev[5]};
    int* info = work.info_ptr();
    LRN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    LRN_CUDA(cudaEventRecord(evStart, st));
    LRN_CUDA(cudaStreamWaitEvent(sp, evStart, 0));
    LRN_CUDA(cudaEventRecord(evU, st));
    for (int p = 0; p < npan; p++) {
        const int b = p & 1;
        const int c0 = p * pw, w = (n - c0 < pw) ? (n - c0) : pw, rows = n - c0, owner = p % ctx.world;
        double* buf = panelbuf.p + (size_t)b * bufsz;
        double* Ap = A + (size_t)c0 * lda + c0;
        double* dk = work.dinv.p + (size_t)(c0 / CHOL_DB) * CHOL_DB * CHOL_DB;
        const int nd = (int)cdiv(w, CHOL_DB) * CHOL_DB * CHOL_DB;
        double* dbuf = buf + (size_t)ldp * pw;
        if (p >= 2) LRN_CUDA(cudaStreamWaitEvent(sp, evE[b], 0));      // buffer b was last read by the updates of step p-2
        if (ctx.rank == owner) {
            LRN_CUDA(cudaStreamWaitEvent(sp, evU, 0));                    // panel p has received the update of step p-1
            cholesky_panel(Ap, rows, w, lda, dk, info, c0, work, buf, ldp, sp);   // factor + pack
            LRN_CUDA(cudaMemcpyAsync(dbuf, dk, (size_t)nd * sizeof(double), cudaMemcpyDeviceToDevice, sp));
        }
        // one broadcast carries the panel and its inverse diagonal blocks (contiguous in buf)
        LRN_NCCL(nccl_api().Broadcast(buf, buf, (size_t)ldp * pw + nd, ncclDouble, owner, ctx.comm, sp));
        LRN_CUDA(cudaEventRecord(evB[b], sp));
        LRN_CUDA(cudaStreamWaitEvent(st, evB[b], 0));
        if (ctx.rank != owner) {                                          // every rank keeps the complete factor
            dim3 grid((unsigned)cdiv(rows, 256), (unsigned)w);
            k_pack_panel<<<grid, 256, 0, st>>>(Ap, lda, rows, w, buf, ldp, 1);
            LRN_CHECK_LAUNCH();
            LRN_CUDA(cudaMemcpyAsync(dk, dbuf, (size_t)nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        auto update = [&](int q) {
            const int q0 = q * pw, wq = (n - q0 < pw) ? (n - q0) : pw;
            const double* Pq = buf + (q0 - c0);
            gemm_nt(st, n - q0, wq, w, -1.0, Pq, ldp, Pq, ldp, 1.0, A + (size_t)q0 * lda + q0, lda);
        };
        if (p + 1 < npan && (p + 1) % ctx.world == ctx.rank) {            // next panel first, so that its owner can go on
            update(p + 1);
            LRN_CUDA(cudaEventRecord(evU, st));
        }
        for (int q = p + 2; q < npan; q++)
            if (q % ctx.world == ctx.rank) update(q);
        LRN_CUDA(cudaEventRecord(evE[b], st));
    }
    // the first failing pivot index is known to the owner of that panel only: take the smallest non-zero over ranks
    int* all = nullptr;
    LRN_CUDA(cudaMalloc(&all, sizeof(int) * ctx.world));
    LRN_NCCL(nccl_api().AllGather(info, all, 1, ncclInt32, ctx.comm, st));
    for (int r = 0; r < ctx.world; r++) k_max_int<<<1, 1, 0, st>>>(info, all + r);
    LRN_CUDA(cudaStreamSynchronize(st));
    cudaFree(all);
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
