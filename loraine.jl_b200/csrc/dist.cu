#include "dist.cuh"
#include "solver.cuh"
#include "chol.cuh"
#include "gemm.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <vector>
#include <unistd.h>
#include <mutex>

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
            api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
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

DistCtx::~DistCtx() {
    for (void* m : ipc_opened) cudaIpcCloseMemHandle(m);
    ipc_opened.clear();
    if (comm && nccl_api().CommDestroy) nccl_api().CommDestroy(comm);
    comm = nullptr;
}

namespace {
// recv[r][z] (pw x pw, ld pw) holds row block g = first_r + z * world of column panel p as solved by rank r; store every
// block at its place in L.  grid = (row chunks, w columns, world * maxcnt slots)
__global__ void k_unpack_blocks(const double* __restrict__ recv, double* __restrict__ L, int lda, int n, int c0, int w, int pw, int p,
                                int world, int maxcnt, int nblk) {
    const int slot = blockIdx.z, r = slot / maxcnt, z = slot - r * maxcnt;
    const int first = p + ((r - p % world + world) % world);
    const int g = first + z * world;
    if (g >= nblk) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int row = g * pw + i;
    if (i >= pw || row >= n || j >= w) return;
    L[(size_t)(c0 + j) * lda + row] = recv[((size_t)slot * pw + j) * pw + i];
}
__global__ void k_min_nonzero(int* a, const int* b, int cnt) {
    int best = 0;
    for (int r = 0; r < cnt; r++) if (b[r] != 0 && (best == 0 || b[r] < best)) best = b[r];
    *a = best;
}
__global__ void k_copy_block(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int rows, int cols) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < rows && j < cols) dst[(size_t)j * ldd + i] = src[(size_t)j * lds + i];
}

// ---- peer-memory exchange kernels ------------------------------------------------------------------------------------------
constexpr long long P2P_SPIN_LIMIT = 4000000000LL;     // ~2 s of SM clocks: a lost peer must not hang the GPU

// copy `count` doubles (multiple of 2) from src to the same offset of every peer buffer, then publish `stamp` in every peer's
// flag word `slot` (the last CTA to finish does it, after a system-wide fence)
__global__ void k_push_vec(const double* src, double* const* __restrict__ peers, int world, int self, size_t dst_off, size_t count,
                           int* const* __restrict__ peer_flags, int slot, int stamp, unsigned int* __restrict__ done_ctr) {
    const size_t n2 = count / 2;
    const double2* s2 = reinterpret_cast<const double2*>(src);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = s2[i];
        for (int r = 0; r < world; r++)
            if (r != self) reinterpret_cast<double2*>(peers[r] + dst_off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_ctr, 1u);
        if (prev == gridDim.x - 1) {
            *done_ctr = 0;
            __threadfence_system();
            for (int r = 0; r < world; r++) *reinterpret_cast<volatile int*>(peer_flags[r] + slot) = stamp;
        }
    }
}
// my solved row blocks of panel p (send[z], pw x w, ld pw; block g = first + z * world) -> every peer's L at their final place
__global__ void k_push_blocks(const double* __restrict__ send, double* const* __restrict__ peers, int world, int lda, int n, int c0,
                              int w, int pw, int first, int cnt, int* const* __restrict__ peer_flags, int slot, int stamp,
                              unsigned int* __restrict__ done_ctr) {
    // grid: (row pairs of a block, column, slot z)
    const int z = blockIdx.z, j = blockIdx.y;
    const int g = first + z * world;
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int row = g * pw + i;
    if (z < cnt && j < w && i < pw && row < n) {
        const double* sp_ = send + ((size_t)z * pw + j) * pw + i;
        const size_t off = (size_t)(c0 + j) * lda + row;
        if (row + 1 < n && i + 1 < pw) {
            const double2 v = *reinterpret_cast<const double2*>(sp_);
            for (int r = 0; r < world; r++) *reinterpret_cast<double2*>(peers[r] + off) = v;
        } else {
            const double v = *sp_;
            for (int r = 0; r < world; r++) peers[r][off] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned int prev = atomicAdd(done_ctr, 1u);
        if (prev == total - 1) {
            *done_ctr = 0;
            __threadfence_system();
            for (int r = 0; r < world; r++) *reinterpret_cast<volatile int*>(peer_flags[r] + slot) = stamp;
        }
    }
}
// wait until flags[slot0 .. slot0 + nslots) have all reached `stamp` (written by the peers over NVLink)
__global__ void k_wait_flags(volatile int* flags, int slot0, int nslots, int stamp) {
    const int t = threadIdx.x;
    if (t < nslots) {
        const long long t0 = clock64();
        while (flags[slot0 + t] < stamp) {
            if (clock64() - t0 > P2P_SPIN_LIMIT) { flags[63] = 1; break; }
        }
    }
    __threadfence_system();
}

__global__ void k_check_timeout(int* flags, int* info) {
    if (flags[63] != 0) { *info = -77; flags[63] = 0; }
}

struct IpcPacket {
    cudaIpcMemHandle_t hx, hl;
    unsigned long long px, pl;          // raw pointers (same-process peers)
    unsigned long long ox, ol;          // byte offsets of the buffers inside their underlying allocations (IPC maps the base)
    long long pid;
    int dev, pad;
};

// offset of a device pointer inside the allocation cudaIpcGetMemHandle exports (cuMemGetAddressRange, resolved at run time)
size_t alloc_offset(const void* ptr) {
    typedef int (*fn_t)(unsigned long long*, size_t*, unsigned long long);
    static fn_t fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* hnd = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (hnd) fn = (fn_t)dlsym(hnd, "cuMemGetAddressRange_v2");
    });
    unsigned long long base = 0;
    size_t size = 0;
    if (fn && fn(&base, &size, (unsigned long long)ptr) == 0 && base) return (size_t)((unsigned long long)ptr - base);
    return 0;
}

// One-time (per factor matrix) exchange of the buffer mappings.  Collective.  Leaves ctx.p2p = 1 on success on ALL ranks,
// 0 otherwise (agreed with a max-reduction so that no rank takes the peer path alone).
void setup_p2p(DistCtx& ctx, double* L, cudaStream_t st) {
    const int world = ctx.world;
    const char* env = getenv("LRN_DIST_P2P");
    int ok = (env && atoi(env) == 0) ? 0 : 1;
    if (world < 2) ok = 0;
    for (void* m : ctx.ipc_opened) cudaIpcCloseMemHandle(m);
    ctx.ipc_opened.clear();
    ctx.flags = reinterpret_cast<int*>(ctx.xb.p + ctx.flags_off);
    IpcPacket mine;
    std::memset(&mine, 0, sizeof mine);
    int dev = 0;
    cudaGetDevice(&dev);
    mine.dev = dev;
    mine.pid = (long long)getpid();
    mine.px = (unsigned long long)ctx.xb.p; mine.pl = (unsigned long long)L;
    mine.ox = alloc_offset(ctx.xb.p); mine.ol = alloc_offset(L);
    if (ok) {
        if (cudaIpcGetMemHandle(&mine.hx, ctx.xb.p) != cudaSuccess || cudaIpcGetMemHandle(&mine.hl, L) != cudaSuccess) {
            cudaGetLastError();
            mine.pid = -1;                       // this rank cannot export: everybody falls back
        }
    }
    static_assert(sizeof(IpcPacket) <= 256, "IpcPacket must fit the exchange slot");
    DevBuf<char> sb(256), rb((size_t)256 * world);
    LRN_CUDA(cudaMemcpyAsync(sb.p, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
    LRN_NCCL(nccl_api().AllGather(sb.p, rb.p, 256, ncclChar, ctx.comm, st));
    std::vector<char> all((size_t)256 * world);
    LRN_CUDA(cudaMemcpyAsync(all.data(), rb.p, all.size(), cudaMemcpyDeviceToHost, st));
    LRN_CUDA(cudaStreamSynchronize(st));
    ctx.peer_xb.assign(world, nullptr); ctx.peer_L.assign(world, nullptr); ctx.peer_flags.assign(world, nullptr);
    for (int r = 0; r < world && ok; r++) {
        IpcPacket pk;
        std::memcpy(&pk, all.data() + (size_t)256 * r, sizeof pk);
        if (pk.pid < 0) { ok = 0; break; }
        if (r == ctx.rank) {
            ctx.peer_xb[r] = ctx.xb.p; ctx.peer_L[r] = L;
        } else if (pk.pid == mine.pid) {         // same process (lrn_create_multi): plain peer access
            int can = 0;
            cudaDeviceCanAccessPeer(&can, dev, pk.dev);
            if (!can) { ok = 0; break; }
            cudaError_t e = cudaDeviceEnablePeerAccess(pk.dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0; break; }
            cudaGetLastError();
            ctx.peer_xb[r] = (double*)pk.px; ctx.peer_L[r] = (double*)pk.pl;
        } else {
            void *mx = nullptr, *ml = nullptr;
            if (cudaIpcOpenMemHandle(&mx, pk.hx, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&ml, pk.hl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                for (void* m : {mx, ml}) if (m) cudaIpcCloseMemHandle(m);
                ok = 0;
                break;
            }
            ctx.ipc_opened.push_back(mx); ctx.ipc_opened.push_back(ml);
            ctx.peer_xb[r] = (double*)((char*)mx + pk.ox); ctx.peer_L[r] = (double*)((char*)ml + pk.ol);
        }
        ctx.peer_flags[r] = reinterpret_cast<int*>(ctx.peer_xb[r] + ctx.flags_off);
    }
    // agree: every rank must have succeeded
    int* dflag = ctx.infos.p;
    const int bad = ok ? 0 : 1;
    LRN_CUDA(cudaMemcpyAsync(dflag, &bad, sizeof(int), cudaMemcpyHostToDevice, st));
    LRN_NCCL(nccl_api().AllReduce(dflag, dflag, 1, ncclInt32, ncclMax, ctx.comm, st));
    int anybad = 1;
    LRN_CUDA(cudaMemcpyAsync(&anybad, dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    LRN_CUDA(cudaStreamSynchronize(st));
    ctx.p2p = anybad ? 0 : ((env && atoi(env) >= 2) ? 2 : 1);
    ctx.L_mapped = L;
    if (ctx.p2p) {
        ctx.d_peer_xb.upload(ctx.peer_xb, st); ctx.d_peer_L.upload(ctx.peer_L, st); ctx.d_peer_flags.upload(ctx.peer_flags, st);
        LRN_CUDA(cudaMemsetAsync(ctx.flags, 0, 66 * sizeof(int), st));
        LRN_CUDA(cudaStreamSynchronize(st));
        // nobody may push before everybody has cleared its flags
        LRN_NCCL(nccl_api().AllReduce(dflag, dflag, 1, ncclInt32, ncclMax, ctx.comm, st));
        LRN_CUDA(cudaStreamSynchronize(st));
        ctx.epoch = 0;
    }
    if (getenv("LRN_DIST_TRACE") && ctx.rank == 0)
        fprintf(stderr, "[lrn dist] exchange mode %d: %s\n", ctx.p2p,
                ctx.p2p == 2 ? "inverse and solved panel pushed over peer memory" :
                ctx.p2p == 1 ? "inverse pushed over peer memory, solved panel through ncclAllGather" : "NCCL broadcast / all-gather");
}

// first row block >= p owned by `rank`, and how many of its blocks are >= p
inline int first_block(int p, int rank, int world) { return p + ((rank - p % world + world) % world); }
inline int count_blocks(int p, int rank, int world, int nblk) {
    const int f = first_block(p, rank, world);
    return f >= nblk ? 0 : (nblk - 1 - f) / world + 1;
}

// C[rows of my blocks g >= gmin, cols [cbeg, cend)] (+)= alpha * f(A[rows, 0:K] * Bm[cbeg:cend, 0:K]^T); `lower` clips every row
// block at its own diagonal block (staircase).  Whole blocks go through one strided-batch launch, a partial last block through
// a second one.
void row_block_gemm(cudaStream_t st, int n, int pw, int rank, int world, int gmin, const double* A, int lda, const double* Bm,
                    int ldb, double* C, int ldc, int cbeg, int cend, int K, double alpha, double beta, int mode, bool lower) {
    const int nblk = (int)cdiv(n, pw);
    const int f = first_block(gmin, rank, world);
    if (f >= nblk || cend <= cbeg) return;
    const int cnt = (nblk - 1 - f) / world + 1;
    const int glast = f + (cnt - 1) * world;
    const bool last_partial = (glast == nblk - 1) && (n - glast * pw < pw);
    const int nfull = last_partial ? cnt - 1 : cnt;
    auto base = [&](GemmParams& g) {
        g.B = Bm + cbeg; g.ldb = ldb; g.lda = lda; g.ldc = ldc; g.K = K; g.transB = true;
        g.alpha = alpha; g.beta = beta; g.mode = mode; g.lower = lower ? 1 : 0; g.col0 = cbeg;
    };
    if (nfull > 0) {
        GemmParams g;
        base(g);
        const int gl = f + (nfull - 1) * world;
        g.A = A + (size_t)f * pw; g.C = C + (size_t)cbeg * ldc + (size_t)f * pw;
        g.M = pw; g.N = (lower ? std::min(cend, (gl + 1) * pw) : cend) - cbeg;
        g.batch = nfull; g.sA = (long long)world * pw; g.sC = (long long)world * pw; g.sB = 0;
        g.row0 = f * pw; g.row0z = (long long)world * pw;
        if (g.N > 0) gemm(g, st);
    }
    if (last_partial) {
        GemmParams g;
        base(g);
        g.A = A + (size_t)glast * pw; g.C = C + (size_t)cbeg * ldc + (size_t)glast * pw;
        g.M = n - glast * pw; g.N = cend - cbeg; g.row0 = glast * pw;
        gemm(g, st);
    }
}
}  // namespace

void dist_allreduce_sum(double* buf, size_t count, DistCtx& ctx, cudaStream_t st) {
    LRN_NCCL(nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, ctx.comm, st));
}

void syrk_sq_row_blocks(const double* BG, int ldbg, int n, int K, double* H, int ldh, int rank, int world, int pw, cudaStream_t st) {
    row_block_gemm(st, n, pw, rank, world, 0, BG, ldbg, BG, ldbg, H, ldh, 0, n, K, 1.0, 1.0, 1, true);
}

void cholesky_dist(double* A, int n, int lda, CholWork& work, DistCtx& ctx, int pw, cudaStream_t st) {
    work.tinv_for = nullptr;   // the triangular solves rebuild their diagonal-block inverses from the new factor
    work.ensure(n);
    LRN_REQUIRE(pw % CHOL_DB == 0 && pw <= 512, "row block height must be a multiple of 64, at most 512");
    LRN_REQUIRE(ctx.comm, "no NCCL communicator (lrn_dist_init)");
    const int world = ctx.world, rank = ctx.rank;
    const int nblk = (int)cdiv(n, pw);
    const int maxcnt0 = (int)cdiv(nblk, world);
    const size_t blk = (size_t)pw * pw, ndmax = (size_t)(pw / CHOL_DB) * CHOL_DB * CHOL_DB;
    const size_t xhalf = 2 * blk + ndmax;            // one inverse buffer: X | 64 x 64 inverse blocks | scratch of the inversion
    // two inverse buffers (step parity: X(p+1) is produced while X(p) is still being used) + flag words of the peer exchange
    if (ctx.xb.n < 2 * xhalf + 64) { ctx.xb.alloc(2 * xhalf + 64); ctx.flags_off = 2 * xhalf; ctx.p2p = -1; }
    if (ctx.sendbuf.n < (size_t)maxcnt0 * blk) ctx.sendbuf.alloc((size_t)maxcnt0 * blk);
    if (ctx.recvbuf.n < (size_t)maxcnt0 * blk * world) ctx.recvbuf.alloc((size_t)maxcnt0 * blk * world);
    if (ctx.infos.n < (size_t)world) ctx.infos.alloc(world);
    if (ctx.p2p < 0 || (ctx.p2p >= 1 && ctx.L_mapped != A)) setup_p2p(ctx, A, st);     // collective, once per factor matrix
    // exchange mode: 0 NCCL only; 1 (default) the 2 MB inverse travels over peer memory (latency-bound), the solved panel through
    // ncclAllGather (bandwidth-bound: NCCL replicates inside the NVSwitch, a peer push sends world-1 copies); 2 everything pushed
    const int mode = ctx.p2p;
    const bool push_x = mode >= 1, push_panel = mode >= 2;
    const int stamp0 = (int)(ctx.epoch * (nblk + 1));
    unsigned int* ctr0 = reinterpret_cast<unsigned int*>(ctx.flags ? ctx.flags + 64 : nullptr);
    unsigned int* ctr1 = reinterpret_cast<unsigned int*>(ctx.flags ? ctx.flags + 65 : nullptr);
    ensure_aux(work);
    cudaStream_t sp = work.aux;    // panel stream: X arrival, row solves, exchange of the solved panel
    cudaStream_t sq = work.aux2;   // diagonal stream: early update + factorisation + inversion of the NEXT diagonal block
    cudaEvent_t evStart = work.ev[0], evU = work.ev[1], evB = work.ev[2], evS = work.ev[3], evD = work.ev[4], evU2 = work.ev[5];
    int* info = work.info_ptr();
    LRN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    LRN_CUDA(cudaEventRecord(evStart, st));
    LRN_CUDA(cudaStreamWaitEvent(sp, evStart, 0));
    LRN_CUDA(cudaStreamWaitEvent(sq, evStart, 0));
    // optional timeline (LRN_DIST_TRACE=1): CUDA events around every stage of the panel chain and of the update, summed per stage
    static const bool trace = getenv("LRN_DIST_TRACE") != nullptr;
    struct Ev { cudaEvent_t e; int stage; };
    std::vector<Ev> tev;
    auto mark = [&](cudaStream_t s_, int stage) {
        if (!trace) return;
        Ev v; v.stage = stage;
        LRN_CUDA(cudaEventCreate(&v.e));
        LRN_CUDA(cudaEventRecord(v.e, s_));
        tev.push_back(v);
    };
    auto xbuf = [&](int p) { return ctx.xb.p + (size_t)(p & 1) * xhalf; };
    // factor diagonal block q (already fully updated), build the inverse of its factor and hand it to the other ranks
    auto factor_and_publish = [&](int q) {
        const int c0 = q * pw, w = std::min(pw, n - c0);
        double* X = xbuf(q);
        double* dk = work.dinv.p + (size_t)(c0 / CHOL_DB) * CHOL_DB * CHOL_DB;
        const int nd = (int)cdiv(w, CHOL_DB) * CHOL_DB * CHOL_DB;
        chol_diag_block(A + (size_t)c0 * lda + c0, lda, w, dk, X, pw, X + blk + ndmax, info, c0, sq);
        LRN_CUDA(cudaMemcpyAsync(X + blk, dk, (size_t)nd * sizeof(double), cudaMemcpyDeviceToDevice, sq));
        if (push_x) {
            k_push_vec<<<96, 256, 0, sq>>>(X, ctx.d_peer_xb.p, world, rank, (size_t)(q & 1) * xhalf, blk + ndmax, ctx.d_peer_flags.p, 0,
                                           stamp0 + q + 1, ctr0);
            LRN_CHECK_LAUNCH();
        }
        LRN_CUDA(cudaEventRecord(evD, sq));
    };
    if (rank == 0 % world) factor_and_publish(0);
    for (int p = 0; p < nblk; p++) {
        const int c0 = p * pw, w = std::min(pw, n - c0), owner = p % world;
        double* X = xbuf(p);
        double* xd = X + blk;
        double* dk = work.dinv.p + (size_t)(c0 / CHOL_DB) * CHOL_DB * CHOL_DB;
        const int nd = (int)cdiv(w, CHOL_DB) * CHOL_DB * CHOL_DB;
        const int stamp = stamp0 + p + 1;
        // ---- panel chain of step p ------------------------------------------------------------------------------------
        if (p > 0) LRN_CUDA(cudaStreamWaitEvent(sp, evU, 0));             // my rows of column block p have the update of step p-1
        mark(sp, 0);
        if (rank == owner) LRN_CUDA(cudaStreamWaitEvent(sp, evD, 0));     // (my own factorisation of block p, on the diagonal stream)
        mark(sp, 1);
        if (push_x) {
            if (rank != owner) {                                          // the owner pushed X(p) into xbuf(p) and raised flag 0
                k_wait_flags<<<1, 32, 0, sp>>>(ctx.flags, 0, 1, stamp);
                LRN_CHECK_LAUNCH();
            }
        } else if (world > 1) {
            LRN_NCCL(nccl_api().Broadcast(X, X, blk + ndmax, ncclDouble, owner, ctx.comm, sp));
        }
        if (rank != owner) LRN_CUDA(cudaMemcpyAsync(dk, xd, (size_t)nd * sizeof(double), cudaMemcpyDeviceToDevice, sp));
        mark(sp, 2);
        const int maxcnt = (int)cdiv(nblk - p, world);
        const int f = first_block(p, rank, world), cnt = count_blocks(p, rank, world, nblk);
        if (p + 1 < nblk || world > 1) {
            // my row blocks g > p of the panel:  send[z] = A[g rows, panel p] * X^T ; the owner's slot 0 is the diagonal block itself
            const int z0 = (rank == owner) ? 1 : 0;
            if (rank == owner) {
                dim3 grid((unsigned)cdiv(w, 256), (unsigned)w);
                k_copy_block<<<grid, 256, 0, sp>>>(A + (size_t)c0 * lda + c0, lda, ctx.sendbuf.p, pw, w, w);
                LRN_CHECK_LAUNCH();
            }
            if (cnt > z0) {
                const int glast = f + (cnt - 1) * world;
                const bool last_partial = (glast == nblk - 1) && (n - glast * pw < pw);
                const int nfull = (last_partial ? cnt - 1 : cnt) - z0;
                if (nfull > 0) {
                    GemmParams g;
                    g.A = A + (size_t)c0 * lda + (size_t)(f + z0 * world) * pw; g.lda = lda; g.sA = (long long)world * pw;
                    g.B = X; g.ldb = pw; g.transB = true;
                    g.C = ctx.sendbuf.p + (size_t)z0 * blk; g.ldc = pw; g.sC = (long long)blk;
                    g.M = pw; g.N = w; g.K = w; g.batch = nfull;
                    gemm(g, sp);
                }
                if (last_partial && glast > p) {
                    GemmParams g;
                    g.A = A + (size_t)c0 * lda + (size_t)glast * pw; g.lda = lda;
                    g.B = X; g.ldb = pw; g.transB = true;
                    g.C = ctx.sendbuf.p + (size_t)(cnt - 1) * blk; g.ldc = pw;
                    g.M = n - glast * pw; g.N = w; g.K = w;
                    gemm(g, sp);
                }
            }
        }
        LRN_CUDA(cudaEventRecord(evS, sp));
        mark(sp, 3);
        // ---- diagonal stream: the owner of block p+1 has just solved ITS OWN rows of panel p, which is all the update of its
        //      diagonal block needs: update, factor, invert and publish X(p+1) while the panel of step p is still being exchanged
        if (p + 1 < nblk && rank == (p + 1) % world) {
            const int c1 = c0 + pw, hb1 = std::min(pw, n - c1);
            LRN_CUDA(cudaStreamWaitEvent(sq, evS, 0));
            if (p > 0) LRN_CUDA(cudaStreamWaitEvent(sq, evU2, 0));        // column block p+1 has the update of step p-1
            const double* Srow = ctx.sendbuf.p + (size_t)((p + 1 - f) / world) * blk;     // my solved block g = p+1
            GemmParams g;
            g.A = Srow; g.B = Srow; g.C = A + (size_t)c1 * lda + c1;
            g.M = hb1; g.N = hb1; g.K = w; g.lda = pw; g.ldb = pw; g.ldc = lda;
            g.transB = true; g.alpha = -1.0; g.beta = 1.0; g.lower = 1;
            gemm(g, sq);
            factor_and_publish(p + 1);
        }
        // ---- exchange of the solved panel -----------------------------------------------------------------------------------
        if (p + 1 < nblk || world > 1) {
            if (push_panel) {
                // every rank stores its solved blocks at their final place in EVERY rank's L (its own included) over NVLink,
                // raises flag 1 + rank everywhere, then waits until the blocks of all ranks have landed here
                dim3 grid((unsigned)cdiv(pw / 2, 128), (unsigned)w, (unsigned)std::max(cnt, 1));
                k_push_blocks<<<grid, 128, 0, sp>>>(ctx.sendbuf.p, ctx.d_peer_L.p, world, lda, n, c0, w, pw, f, cnt, ctx.d_peer_flags.p,
                                                    1 + rank, stamp, ctr1);
                LRN_CHECK_LAUNCH();
                mark(sp, 4);
                k_wait_flags<<<1, 64, 0, sp>>>(ctx.flags, 1, world, stamp);
                LRN_CHECK_LAUNCH();
            } else {
                const double* src = ctx.sendbuf.p;
                if (world > 1) {
                    LRN_NCCL(nccl_api().AllGather(ctx.sendbuf.p, ctx.recvbuf.p, (size_t)maxcnt * blk, ncclDouble, ctx.comm, sp));
                    src = ctx.recvbuf.p;
                }
                mark(sp, 4);
                dim3 grid((unsigned)cdiv(pw, 256), (unsigned)w, (unsigned)(world * maxcnt));
                k_unpack_blocks<<<grid, 256, 0, sp>>>(src, A, lda, n, c0, w, pw, p, world, maxcnt, nblk);
                LRN_CHECK_LAUNCH();
            }
            mark(sp, 5);
        }
        LRN_CUDA(cudaEventRecord(evB, sp));
        LRN_CUDA(cudaStreamWaitEvent(st, evB, 0));
        if (p + 1 >= nblk) break;
        // ---- trailing update on the main stream.  Two column blocks ahead first: column block p+1 (my rows are solved next; its
        //      diagonal block was already updated on the diagonal stream) and column block p+2 (its diagonal block is factored
        //      during step p+1).  The REST of the trailing matrix is updated every second step only, with the two panels p-1, p at
        //      once (K = 2 pw: half as many C-tile read-modify-write epilogues and pipeline fills per flop on the TMA-fed kernel):
        //        even step p : columns p+1, p+2 with panel p                                         (rest deferred)
        //        odd step p  : column p+1 with panel p, column p+2 with panels p-1..p, rest (>= p+3) with panels p-1..p
        const double* P = A + (size_t)c0 * lda;                            // column panel p: rows are global
        const int c1 = c0 + pw, c2 = std::min(n, c1 + pw), c3 = std::min(n, c2 + pw);
        const bool pair_rest = (pw % 32 == 0);                             // (always: pw is 128, 256 or 512)
        mark(st, 10);
        if (!(p & 1) || !pair_rest) {
            // (one launch for both look-ahead column blocks: N = 2 pw keeps the strided-batch product on the TMA-fed kernel)
            row_block_gemm(st, n, pw, rank, world, p + 2, P, lda, P, lda, A, lda, c1, c3, w, -1.0, 1.0, 0, true);
            LRN_CUDA(cudaEventRecord(evU, st));
            LRN_CUDA(cudaEventRecord(evU2, st));
            mark(st, 11);
            const bool last_even_with_rest = pair_rest && (p + 2 >= nblk);   // no odd step follows that could take the rest
            if ((!pair_rest || last_even_with_rest) && c3 < n)
                row_block_gemm(st, n, pw, rank, world, p + 3, P, lda, P, lda, A, lda, c3, n, w, -1.0, 1.0, 0, true);
        } else {
            const double* P2 = P - (size_t)pw * lda;                       // panels p-1 and p side by side: K = 2 pw
            row_block_gemm(st, n, pw, rank, world, p + 2, P, lda, P, lda, A, lda, c1, c2, w, -1.0, 1.0, 0, true);
            LRN_CUDA(cudaEventRecord(evU, st));
            if (c2 < n) row_block_gemm(st, n, pw, rank, world, p + 2, P2, lda, P2, lda, A, lda, c2, c3, pw + w, -1.0, 1.0, 0, true);
            LRN_CUDA(cudaEventRecord(evU2, st));
            mark(st, 11);
            if (c3 < n) row_block_gemm(st, n, pw, rank, world, p + 3, P2, lda, P2, lda, A, lda, c3, n, pw + w, -1.0, 1.0, 0, true);
        }
        mark(st, 12);
    }
    LRN_CUDA(cudaEventRecord(evD, sq));
    LRN_CUDA(cudaStreamWaitEvent(st, evD, 0));
    if (trace) {
        LRN_CUDA(cudaStreamSynchronize(sp));
        LRN_CUDA(cudaStreamSynchronize(sq));
        LRN_CUDA(cudaStreamSynchronize(st));
        // stage sums on the panel stream: 0->1 wait for my own diagonal block, 1->2 X arrival, 2->3 row solves, 3->4 exchange,
        // 4->5 unpack / flag wait, 5->next 0 wait for the update of the main stream
        double sum[16] = {0};
        for (size_t i = 0; i + 1 < tev.size(); i++) {
            const int a = tev[i].stage, b2 = tev[i + 1].stage;
            float ms = 0.f;
            if (b2 == a + 1 && a != 5) { cudaEventElapsedTime(&ms, tev[i].e, tev[i + 1].e); sum[a] += ms; }
        }
        const Ev* last5 = nullptr;
        for (auto& v : tev) {
            if (v.stage == 5) last5 = &v;
            if (v.stage == 0 && last5) { float ms = 0.f; cudaEventElapsedTime(&ms, last5->e, v.e); sum[6] += ms; last5 = nullptr; }
        }
        float total = 0.f;
        if (!tev.empty()) cudaEventElapsedTime(&total, tev.front().e, tev.back().e);
        if (rank == 0)
            fprintf(stderr, "[lrn dist trace] world %d n %d mode %d: own-diag wait %.2f  X arrival %.2f  solve %.2f  exchange %.2f  unpack/flags %.2f  "
                            "panel-stream wait %.2f | 2 look-ahead column blocks %.2f  rest update %.2f | first-to-last event %.2f ms\n",
                    world, n, mode, sum[0], sum[1], sum[2], sum[3], sum[4], sum[6], sum[10], sum[11], total);
        for (auto& v : tev) cudaEventDestroy(v.e);
    }
    ctx.epoch++;
    if (push_x) {                                  // a peer that never answered: report instead of returning a wrong factor
        k_check_timeout<<<1, 1, 0, st>>>(ctx.flags, info);
        LRN_CHECK_LAUNCH();
    }
    // the first failing pivot index is known to the owner of that block only: take the smallest non-zero over ranks
    if (world > 1) {
        LRN_NCCL(nccl_api().AllGather(info, ctx.infos.p, 1, ncclInt32, ctx.comm, st));
        k_min_nonzero<<<1, 1, 0, st>>>(info, ctx.infos.p, world);
        LRN_CHECK_LAUNCH();
    }
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
    if (h->group) { h->err = "lrn_dist_init: the handle already drives several GPUs in-process (lrn_create_multi)"; return LRN_ERR_STATE; }
    try {
        LRN_CUDA(cudaSetDevice(h->device));
        ncclUniqueId id;
        std::memcpy(&id, unique_id128, sizeof id);
        auto* ctx = new DistCtx();
        ctx->rank = rank;
        ctx->world = world;
        LRN_NCCL(nccl_api().CommInitRank(&ctx->comm, world, id, rank));
        if (h->nccl) delete static_cast<DistCtx*>(h->nccl);
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
