// Blocked right-looking Cholesky: 64x64 diagonal blocks are factored AND inverted by one CTA in shared memory; the panel
// TRSM is two DMMA GEMMs per 64-column block (using the inverted diagonal blocks); the trailing update is a lower-only
// DMMA SYRK-shaped GEMM with K = panel width (64..512, recursion keeps the bulk of the flops in large-K updates).
#include "chol.cuh"
#include "gemm.cuh"

namespace lrn {
namespace {

constexpr int DB = CHOL_DB;
constexpr int SLD = DB + 1;
constexpr size_t POTRF_SMEM = (2 * DB * SLD + DB + 2 + 4 * DB) * sizeof(double);

// One CTA factors a 64x64 diagonal block held in REGISTERS (each of the 256 threads owns a cyclic 4x4 sub-tile: rows
// ti+16a, columns tj+16b), two barriers per column; the inverse of the factor is then built row by row in shared memory
// (4 partial dot products per entry).  ~15 us instead of ~100 us for the previous shared-memory version.
__global__ void __launch_bounds__(256)
    potrf_diag_kernel(double* __restrict__ A, int lda, int nb, double* __restrict__ dinv, int* __restrict__ info, int base) {
    extern __shared__ double sm[];
    double* sL = sm;                      // factor (lower), later read by the inversion   [DB][SLD]
    double* sX = sm + DB * SLD;           // inverse                                        [DB][SLD]
    double* colbuf = sX + DB * SLD;       // [DB]
    double* dbuf = colbuf + DB;           // [2]
    double (*part)[DB] = reinterpret_cast<double (*)[DB]>(dbuf + 2);   // [4][DB]
    const int tid = threadIdx.x, ti = tid & 15, tj = tid >> 4;
    double r[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int i = ti + 16 * a, k = tj + 16 * b;
            double v = 0.0;
            if (i < nb && k < nb) v = (i >= k) ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k];   // symmetric fill from the lower part
            else if (i == k) v = 1.0;                                                            // identity padding
            r[a][b] = v;
        }
    for (int j = 0; j < DB; j++) {
        const int ja = j >> 4, jm = j & 15;
        if (ti == jm && tj == jm) {                       // owner of the diagonal entry
            double d = 0.0;
#pragma unroll
            for (int a = 0; a < 4; a++) if (a == ja) d = r[a][a];
            if (!(d > 0.0)) {
                if (j < nb && *info == 0) *info = base + j + 1;
                d = 1.0;
            }
            dbuf[j & 1] = sqrt(d);
        }
        __syncthreads();
        const double djj = dbuf[j & 1];
        if (tj == jm) {                                   // owners of column j publish l(:,j)
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int i = ti + 16 * a;
                double v = 0.0;
#pragma unroll
                for (int b = 0; b < 4; b++) if (b == ja) v = r[a][b];
                v = (i > j) ? v / djj : (i == j ? djj : 0.0);
                colbuf[i] = v;
                sL[i * SLD + j] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = ti + 16 * a;
            const double li = colbuf[i];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int k = tj + 16 * b;
                if (i > j && k > j) r[a][b] -= li * colbuf[k];
            }
        }
        // colbuf of step j is re-written only after the next barrier pair; dbuf is double buffered
    }
    __syncthreads();
    // inverse, row by row: x(i,c) = -( sum_{k=c}^{i-1} L(i,k) x(k,c) ) / L(i,i),  x(i,i) = 1 / L(i,i)
    const int c = tid & 63, pr = tid >> 6;
    for (int i = 0; i < DB; i++) {
        double acc = 0.0;
        for (int k = c + pr; k < i; k += 4) acc += sL[i * SLD + k] * sX[k * SLD + c];
        part[pr][c] = acc;
        __syncthreads();
        if (pr == 0) {
            double v = 0.0;
            const double lii = sL[i * SLD + i];
            if (c == i) v = 1.0 / lii;
            else if (c < i) v = -(part[0][c] + part[1][c] + part[2][c] + part[3][c]) / lii;
            sX[i * SLD + c] = v;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < DB * DB; idx += 256) {
        int i = idx % DB, j = idx / DB;
        if (i < nb && j < nb && i >= j) A[(size_t)j * lda + i] = sL[i * SLD + j];
        dinv[(size_t)j * DB + i] = (i < nb && j < nb) ? sX[i * SLD + j] : 0.0;
    }
}

int pick_nb(int n) {
    if (n > 8192) return 512;
    if (n > 2048) return 256;
    if (n > 512) return 128;
    return 64;
}

// P (rows x kb, below a factored kb x kb diagonal block Akk) <- P * inv(L_kk)^T, 64 columns at a time
void panel_trsm(double* P, int rows, int kb, const double* Akk, const double* dk, int lda, cudaStream_t st) {
    for (int j = 0; j < kb; j += DB) {
        const int jb = (kb - j < DB) ? (kb - j) : DB;
        double* Pj = P + (size_t)j * lda;
        // Pj -= P[:,0:j] * L[k+j : k+j+jb, k : k+j]^T
        if (j > 0) gemm_nt(st, rows, jb, j, -1.0, P, lda, Akk + j, lda, 1.0, Pj, lda);
        // Pj <- Pj * inv(L_jj)^T   (in place: every CTA owns its rows and a single N tile)
        gemm_nt(st, rows, jb, jb, 1.0, Pj, lda, dk + (size_t)(j / DB) * DB * DB, DB, 0.0, Pj, lda);
    }
}

__global__ void k_copy2d(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int rows, int cols) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < rows && j < cols) dst[(size_t)j * ldd + i] = src[(size_t)j * lds + i];
}
void copy2d(const double* src, int lds, double* dst, int ldd, int rows, int cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)cdiv(rows, 256), (unsigned)cols);
    k_copy2d<<<grid, 256, 0, st>>>(src, lds, dst, ldd, rows, cols);
    LRN_CHECK_LAUNCH();
}

// X = inv(L_kk) (kb x kb lower, leading dimension ldx) from the factor and its inverted 64x64 diagonal blocks:
// block row i:  X[i, 0:i) = -inv(L_ii) * ( L[i, 0:i) * X[0:i, 0:i) ),  X[i,i] = inv(L_ii).   2 small GEMMs per block row.
void trtri_blocked(const double* Lkk, int kb, int lda, const double* dk, double* X, int ldx, double* T, cudaStream_t st) {
    LRN_CUDA(cudaMemsetAsync(X, 0, (size_t)ldx * kb * sizeof(double), st));
    const int nb = (int)cdiv(kb, DB);
    for (int i = 0; i < nb; i++) {
        const int i0 = i * DB, ib = (kb - i0 < DB) ? (kb - i0) : DB;
        const double* di = dk + (size_t)i * DB * DB;
        copy2d(di, DB, X + (size_t)i0 * ldx + i0, ldx, ib, ib, st);
        if (i > 0) {
            gemm_nn(st, ib, i0, i0, 1.0, Lkk + i0, lda, X, ldx, 0.0, T, DB);
            gemm_nn(st, ib, i0, ib, -1.0, di, DB, T, DB, 0.0, X + i0, ldx);
        }
    }
}

// rows below a factored kb x kb diagonal block: P <- P * inv(L_kk)^T as ONE large GEMM through the explicit inverse
// (out of place into work.pout or the caller's buffer, then copied back)
void panel_trsm_inv(double* P, int rows, int kb, const double* Akk, const double* dk, int lda, CholWork& work, double* Pout,
                    int ldp, cudaStream_t st) {
    const int ldx = pad_ld(kb);
    const size_t need = (size_t)ldx * kb + (size_t)DB * kb;
    if (work.xinv.n < need) work.xinv.alloc(need);
    double* X = work.xinv.p;
    double* T = X + (size_t)ldx * kb;
    trtri_blocked(Akk, kb, lda, dk, X, ldx, T, st);
    double* out = Pout;
    int ldo = ldp;
    if (!out) {
        ldo = pad_ld(rows);
        if (work.pout.n < (size_t)ldo * kb) work.pout.alloc((size_t)ldo * kb);
        out = work.pout.p;
    }
    gemm_nt(st, rows, kb, kb, 1.0, P, lda, X, ldx, 0.0, out, ldo);
    copy2d(out, ldo, P, lda, rows, kb, st);
}

void chol_rec(double* A, int n, int lda, double* dinv, int* info, int base, CholWork& work, cudaStream_t st) {
    if (n <= DB) {
        potrf_diag_kernel<<<1, 256, POTRF_SMEM, st>>>(A, lda, n, dinv, info, base);
        LRN_CHECK_LAUNCH();
        return;
    }
    const int NB = pick_nb(n);
    for (int k = 0; k < n; k += NB) {
        const int kb = (n - k < NB) ? (n - k) : NB;
        double* Akk = A + (size_t)k * lda + k;
        double* dk = dinv + (size_t)(k / DB) * DB * DB;
        chol_rec(Akk, kb, lda, dk, info, base + k, work, st);
        const int rows = n - k - kb;
        if (rows <= 0) break;
        double* P = A + (size_t)k * lda + (k + kb);          // rows x kb panel below the diagonal block
        if (kb >= 128 && rows >= 1024) panel_trsm_inv(P, rows, kb, Akk, dk, lda, work, nullptr, 0, st);
        else panel_trsm(P, rows, kb, Akk, dk, lda, st);
        // trailing update, lower triangle only
        GemmParams p;
        p.A = P; p.B = P; p.C = A + (size_t)(k + kb) * lda + (k + kb);
        p.M = rows; p.N = rows; p.K = kb; p.lda = lda; p.ldb = lda; p.ldc = lda;
        p.transB = true; p.alpha = -1.0; p.beta = 1.0; p.lower = 1;
        gemm(p, st);
    }
}

__global__ void __launch_bounds__(256)
    trsv_fwd_step(const double* __restrict__ L, int lda, int n, int j0, int jb, const double* __restrict__ dinv,
                  double* __restrict__ rhs, double* __restrict__ sol) {
    __shared__ double xj[DB], yj[DB];
    const int tid = threadIdx.x;
    if (tid < jb) xj[tid] = rhs[j0 + tid];
    __syncthreads();
    if (tid < jb) {
        double acc = 0.0;
        for (int c = 0; c <= tid; c++) acc += dinv[(size_t)c * DB + tid] * xj[c];
        yj[tid] = acc;
        if (blockIdx.x == 0) sol[j0 + tid] = acc;
    }
    __syncthreads();
    const int i = j0 + jb + blockIdx.x * 256 + tid;
    if (i < n) {
        double acc = 0.0;
        const double* Lp = L + (size_t)j0 * lda + i;
#pragma unroll 8
        for (int c = 0; c < jb; c++) acc += Lp[(size_t)c * lda] * yj[c];
        rhs[i] -= acc;
    }
}

__global__ void __launch_bounds__(256)
    trsv_bwd_step(const double* __restrict__ L, int lda, int n, int j0, int jb, const double* __restrict__ dinv,
                  double* __restrict__ rhs, double* __restrict__ sol) {
    __shared__ double xj[DB], yj[DB];
    const int tid = threadIdx.x;
    if (tid < jb) xj[tid] = rhs[j0 + tid];
    __syncthreads();
    if (tid < jb) {
        double acc = 0.0;
        for (int c = tid; c < jb; c++) acc += dinv[(size_t)tid * DB + c] * xj[c];   // (dinv^T)[tid][c] = dinv[c][tid]
        yj[tid] = acc;
        if (blockIdx.x == 0) sol[j0 + tid] = acc;
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + tid;
    if (i < j0) {
        double acc = 0.0;
        const double* Lp = L + (size_t)i * lda + j0;
#pragma unroll 8
        for (int c = 0; c < jb; c++) acc += Lp[c] * yj[c];
        rhs[i] -= acc;
    }
}

__global__ void zero_upper_kernel(double* A, int n, int lda) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i < n && i < j) A[(size_t)j * lda + i] = 0.0;
}

}  // namespace

void cholesky_panel(double* Apanel, int rows, int w, int lda, double* dinv, int* info, int base, CholWork& work, double* Pout,
                    int ldp, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        LRN_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM));
        configured = true;
    }
    chol_rec(Apanel, w, lda, dinv, info, base, work, st);
    if (rows > w) {
        if (w >= 128 && rows - w >= 1024) panel_trsm_inv(Apanel + w, rows - w, w, Apanel, dinv, lda, work, Pout ? Pout + w : nullptr, ldp, st);
        else {
            panel_trsm(Apanel + w, rows - w, w, Apanel, dinv, lda, st);
            if (Pout) copy2d(Apanel + w, lda, Pout + w, ldp, rows - w, w, st);
        }
    }
    if (Pout) copy2d(Apanel, lda, Pout, ldp, w, w, st);
}

void cholesky_lower(double* A, int n, int lda, CholWork& work, cudaStream_t st) {
    work.ensure(n);
    static bool configured = false;
    if (!configured) {
        LRN_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM));
        configured = true;
    }
    LRN_CUDA(cudaMemsetAsync(work.info_ptr(), 0, sizeof(int), st));
    if (n <= 0) return;
    chol_rec(A, n, lda, work.dinv.p, work.info_ptr(), 0, work, st);
}

void chol_solve(const double* L, int n, int lda, const CholWork& work, double* x, double* tmp, int which, cudaStream_t st) {
    if (n <= 0) return;
    const int nblk = (int)cdiv(n, DB);
    if (which & 1) {
        for (int b = 0; b < nblk; b++) {
            int j0 = b * DB, jb = (n - j0 < DB) ? (n - j0) : DB;
            int rest = n - j0 - jb;
            int grid = rest > 0 ? (int)cdiv(rest, 256) : 1;
            trsv_fwd_step<<<grid, 256, 0, st>>>(L, lda, n, j0, jb, work.dinv.p + (size_t)b * DB * DB, x, tmp);
            LRN_CHECK_LAUNCH();
        }
        LRN_CUDA(cudaMemcpyAsync(x, tmp, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    if (which & 2) {
        for (int b = nblk - 1; b >= 0; b--) {
            int j0 = b * DB, jb = (n - j0 < DB) ? (n - j0) : DB;
            int grid = j0 > 0 ? (int)cdiv(j0, 256) : 1;
            trsv_bwd_step<<<grid, 256, 0, st>>>(L, lda, n, j0, jb, work.dinv.p + (size_t)b * DB * DB, x, tmp);
            LRN_CHECK_LAUNCH();
        }
        LRN_CUDA(cudaMemcpyAsync(x, tmp, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
}

void trsm_left_lower_trans(const double* L, int n, int lda, const CholWork& work, double* Y, int ldy, int ncols, cudaStream_t st) {
    if (n <= 0 || ncols <= 0) return;
    const int nblk = (int)cdiv(n, DB);
    for (int b = nblk - 1; b >= 0; b--) {
        const int j0 = b * DB, jb = (n - j0 < DB) ? (n - j0) : DB;
        double* Yj = Y + j0;
        // X_j = inv(L_jj)^T Y_j   (in place: one M tile, every CTA reads exactly the columns it writes)
        gemm_tn(st, jb, ncols, jb, 1.0, work.dinv.p + (size_t)b * DB * DB, DB, Yj, ldy, 0.0, Yj, ldy);
        // Y[0:j0, :] -= L[j0:j0+jb, 0:j0]^T X_j
        if (j0 > 0) gemm_tn(st, j0, ncols, jb, -1.0, L + j0, lda, Yj, ldy, 1.0, Y, ldy);
    }
}

void zero_strict_upper(double* A, int n, int lda, cudaStream_t st) {
    if (n <= 1) return;
    dim3 grid((unsigned)cdiv(n, 256), (unsigned)n);
    zero_upper_kernel<<<grid, 256, 0, st>>>(A, n, lda);
    LRN_CHECK_LAUNCH();
}

}  // namespace lrn
