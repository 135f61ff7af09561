// Blocked right-looking Cholesky: 64x64 diagonal blocks are factored AND inverted by one CTA in shared memory; the panel
// TRSM is two DMMA GEMMs per 64-column block (using the inverted diagonal blocks); the trailing update is a lower-only
// DMMA SYRK-shaped GEMM with K = panel width (64..512, recursion keeps the bulk of the flops in large-K updates).
#include "chol.cuh"
#include "gemm.cuh"

namespace lrn {
namespace {

constexpr int DB = CHOL_DB;
constexpr int SLD = DB + 1;

__global__ void __launch_bounds__(256)
    potrf_diag_kernel(double* __restrict__ A, int lda, int nb, double* __restrict__ dinv, int* __restrict__ info, int base) {
    extern __shared__ double sm[];
    double* s = sm;                 // L      [DB][SLD]
    double* x = sm + DB * SLD;      // L^{-1} [DB][SLD]
    const int tid = threadIdx.x;
    for (int idx = tid; idx < DB * DB; idx += 256) {
        int i = idx % DB, j = idx / DB;
        double v = 0.0;
        if (i < nb && j < nb && i >= j) v = A[(size_t)j * lda + i];
        if (i >= nb && i == j) v = 1.0;   // identity padding keeps the inverse well defined
        s[i * SLD + j] = v;
        x[i * SLD + j] = 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; j++) {
        if (tid == 0) {
            double d = s[j * SLD + j];
            if (!(d > 0.0)) {          // also catches NaN
                if (*info == 0) *info = base + j + 1;
                d = 1.0;
            }
            s[j * SLD + j] = sqrt(d);
        }
        __syncthreads();
        const double djj = s[j * SLD + j];
        for (int i = j + 1 + tid; i < nb; i += 256) s[i * SLD + j] /= djj;
        __syncthreads();
        const int t = nb - 1 - j;
        for (int idx = tid; idx < t * t; idx += 256) {
            int ii = idx % t, kk = idx / t;
            if (ii >= kk) {
                int i = j + 1 + ii, k = j + 1 + kk;
                s[i * SLD + k] -= s[i * SLD + j] * s[k * SLD + j];
            }
        }
        __syncthreads();
    }
    // inverse of the lower-triangular factor, one thread per column
    if (tid < DB) {
        const int c = tid;
        x[c * SLD + c] = 1.0 / s[c * SLD + c];
        for (int i = c + 1; i < DB; i++) {
            double acc = 0.0;
            for (int k = c; k < i; k++) acc += s[i * SLD + k] * x[k * SLD + c];
            x[i * SLD + c] = -acc / s[i * SLD + i];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < DB * DB; idx += 256) {
        int i = idx % DB, j = idx / DB;
        if (i < nb && j < nb && i >= j) A[(size_t)j * lda + i] = s[i * SLD + j];
        dinv[(size_t)j * DB + i] = (i < nb && j < nb) ? x[i * SLD + j] : 0.0;
    }
}

int pick_nb(int n) {
    if (n > 8192) return 512;
    if (n > 2048) return 256;
    if (n > 512) return 128;
    return 64;
}

void chol_rec(double* A, int n, int lda, double* dinv, int* info, int base, cudaStream_t st) {
    if (n <= DB) {
        potrf_diag_kernel<<<1, 256, 2 * DB * SLD * sizeof(double), st>>>(A, lda, n, dinv, info, base);
        LRN_CHECK_LAUNCH();
        return;
    }
    const int NB = pick_nb(n);
    for (int k = 0; k < n; k += NB) {
        const int kb = (n - k < NB) ? (n - k) : NB;
        double* Akk = A + (size_t)k * lda + k;
        double* dk = dinv + (size_t)(k / DB) * DB * DB;
        chol_rec(Akk, kb, lda, dk, info, base + k, st);
        const int rows = n - k - kb;
        if (rows <= 0) break;
        double* P = A + (size_t)k * lda + (k + kb);          // rows x kb panel below the diagonal block
        for (int j = 0; j < kb; j += DB) {
            const int jb = (kb - j < DB) ? (kb - j) : DB;
            double* Pj = P + (size_t)j * lda;
            if (j > 0) {
                // Pj -= P[:,0:j] * L[k+j : k+j+jb, k : k+j]^T
                gemm_nt(st, rows, jb, j, -1.0, P, lda, Akk + j, lda, 1.0, Pj, lda);
            }
            // Pj <- Pj * inv(L_jj)^T   (in place: every CTA owns its rows and a single N tile)
            gemm_nt(st, rows, jb, jb, 1.0, Pj, lda, dk + (size_t)(j / DB) * DB * DB, DB, 0.0, Pj, lda);
        }
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

void cholesky_lower(double* A, int n, int lda, CholWork& work, cudaStream_t st) {
    work.ensure(n);
    static bool configured = false;
    if (!configured) {
        LRN_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(2 * DB * SLD * sizeof(double))));
        configured = true;
    }
    LRN_CUDA(cudaMemsetAsync(work.info_ptr(), 0, sizeof(int), st));
    if (n <= 0) return;
    chol_rec(A, n, lda, work.dinv.p, work.info_ptr(), 0, st);
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

void zero_strict_upper(double* A, int n, int lda, cudaStream_t st) {
    if (n <= 1) return;
    dim3 grid((unsigned)cdiv(n, 256), (unsigned)n);
    zero_upper_kernel<<<grid, 256, 0, st>>>(A, n, lda);
    LRN_CHECK_LAUNCH();
}

}  // namespace lrn
