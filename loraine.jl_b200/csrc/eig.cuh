// Symmetric eigen / SVD machinery of the NT-scaling and step-length code paths.
//   * jacobi_eig_small : batched two-sided Jacobi eigensolver for symmetric matrices up to 64x64 (one CTA each, smem).
//   * svd_block_jacobi : one-sided (Hestenes) block Jacobi SVD of a dense m x m matrix; column-block pairs are rotated
//                        through their 64x64 Gram matrices (only the 32x32 cross block is recomputed per round, the
//                        diagonal blocks travel with the column blocks); all O(m^3) work is DMMA (cross Gram GEMM,
//                        persistent TMA-fed panel rotation).  Replaces FameSVD.fsvd at src/prepare_W.jl:42.
//   * lanczos_extreme  : extreme eigenpairs of a dense symmetric matrix by Lanczos with full re-orthogonalisation.
//                        Replaces `eigmin` (src/predictor_corrector.jl:272,285; src/Solvers.jl:503,505) and the
//                        `eigen(W)` calls of the preconditioners (src/Solvers.jl:642,706), which only use the erank
//                        largest eigenpairs, the smallest eigenvalue and the trace.
#pragma once
#include "common.cuh"

namespace lrn {

struct EigSmallParams {
    const double* A = nullptr;   // batch of symmetric n x n (full storage), element stride sA; sum of `nparts` partials (stride sPart)
    int lda = 0;
    long long sA = 0;
    int nparts = 1;
    long long sPart = 0;
    int n = 0;                   // <= 64
    double* evals = nullptr;     // optional, n per batch (stride sE); sorted descending if sort_desc
    long long sE = 0;
    double* V = nullptr;         // optional eigenvectors (columns), ldv, stride sV
    int ldv = 0;
    long long sV = 0;
    int sort_desc = 0;
    int relative = 0;            // 1: PD/Gram input, purely relative rotation threshold (high relative accuracy)
    double* offmax = nullptr;    // optional: atomicMax of max_{i<j} |a_ij|/sqrt(|a_ii a_jj|) of the INPUT matrices
    double* minval = nullptr;    // optional: per batch smallest eigenvalue (stride 1)
    int batch = 1;
    int max_sweeps = 40;         // cyclic sweeps over all pairs (1 = a single sweep, used inside the block-Jacobi SVD)
    int cross_only = 0;          // n = 64 only: rotate just the 32 x 32 pairs (p < 32 <= q) in 32 steps -- the pairs inside each
                                 // half are covered once per outer sweep by the round that runs the full schedule
    // block-Jacobi SVD bookkeeping (n = 64): the two diagonal 32 x 32 blocks of the rotated matrix J'AJ are the Gram
    // matrices of the two rotated column blocks; they are handed to the pairs of the next round through a per-slot store
    // so that only the 32 x 32 cross block has to be recomputed from the columns.
    double* dg_out = nullptr;        // nslots * 1024 (slot = slotmap[2z], slotmap[2z+1])
    const int* slotmap = nullptr;
    const double* dg_in = nullptr;   // non-null: A holds only the cross block B_{2z+1}' B_{2z} (32 x 32, lda, partials);
                                     // the diagonal blocks come from dg_in[2z], dg_in[2z+1]
};
void jacobi_eig_small(const EigSmallParams& p, cudaStream_t st);

// One block-Jacobi sweep is a fixed sequence of (mp/32 - 1) rounds x 3 small kernels whose operands only depend on the parity
// of the sweep (the ping-pong work buffers swap once per round): at m <= ~2000 the sweep is bound by launch latency, so it is
// captured ONCE per workspace as a CUDA graph (one per parity) and replayed for every later sweep of every later call.
struct SweepGraphs {
    cudaGraphExec_t exec[2] = {nullptr, nullptr};
    long long nodes[2] = {0, 0};
    bool warm = false;           // one eager sweep has run on this workspace (lazy kernel configuration is done)
    bool broken = false;         // capture failed once: stay eager
    void reset();
    ~SweepGraphs() { reset(); }
};

struct SvdWork {
    SweepGraphs graphs;
    int m = 0, mp = 0;           // mp = m padded to a multiple of 64
    DevBuf<double> buf0, buf1;   // (m + mp) x mp stacked [A; V] ping-pong buffers, ld = ldw
    int ldw = 0;
    DevBuf<double> gram;         // pairs * splits * 64*64
    DevBuf<double> rot;          // pairs * 64*64
    DevBuf<double> offmax;       // 1
    DevBuf<int> slotmap;         // mp/32 : round-robin slot permutation
    DevBuf<double> sv;           // mp singular values (unsorted)
    DevBuf<int> perm;            // mp
    DevBuf<double> dg0, dg1;     // per-slot 32 x 32 Gram blocks handed from round to round (ping-pong)
    int splits = 1, Kc = 0;      // split-K of the full 64 x 64 Gram products (first round of a sweep)
    int xsplits = 1, xKc = 0;    // split-K of the 32 x 32 cross products (all other rounds)
    int inner_sweeps = 1;
    bool want_V = true;
    bool panel = false;          // rows padded to 128-row tiles, panel_rotate kernel for the updates
    void ensure(int m_, bool want_V_ = true);
};

// On return (asynchronous): `U_D` (m x m, ldu) holds A*V = U*diag(sigma) with columns sorted by sigma descending,
// `V` (m x m, ldv) the right singular vectors in the same order, `sigma` (m) the singular values.
// V may be null: the right singular vectors are then not accumulated (one third fewer flops).
// Returns the number of sweeps used (host-synchronises once per sweep to read the convergence measure).
int svd_block_jacobi(const double* A, int lda, int m, double* U_D, int ldu, double* V, int ldv, double* sigma, SvdWork& w,
                     double tol, int max_sweeps, cudaStream_t st);

// Batched variant for `nb` matrices of the SAME size m (multi-block SDPs): all blocks advance through the tournament in
// lock step, so one Gram GEMM / one rotation kernel / one update GEMM per round serve every block.  No V accumulation.
struct SvdBatchWork {
    SweepGraphs graphs;
    int m = 0, mp = 0, ldw = 0, nb = 0, splits = 1, Kc = 0;
    DevBuf<double> buf0, buf1, gram, rot, offmax, sv;
    DevBuf<int> slotmap, perm;
    void ensure(int m_, int nb_);
};
int svd_block_jacobi_batched(const double* const* A, int lda, int m, int nb, double* const* UD, int ldu, double* const* sigma,
                             SvdBatchWork& w, double tol, int max_sweeps, cudaStream_t st);

struct LanczosWork {
    int m = 0, kmax = 0;
    DevBuf<double> Q;            // m x (kmax+1)
    DevBuf<double> w, c;         // m, kmax+1
    DevBuf<double> scal;         // small device scalars
    DevBuf<double> ab;           // 2 * kmax: Lanczos alpha_j, beta_j (read back at the check points only)
    double* h_scal = nullptr;    // pinned host mirror
    DevBuf<double> S;            // kmax x nev Ritz coefficient upload
    void ensure(int m_, int kmax_);
    ~LanczosWork();
};

struct LanczosResult {
    double lmin = 0, lmax = 0;
    double resid_min = 0, resid_max = 0;   // Ritz residual bounds |beta_k z| of the two extreme Ritz pairs
    int iters = 0;
    bool converged = false;
};

// Extreme eigenvalues of the symmetric matrix T (m x m, full storage, ld). If nev_top > 0 also returns the nev_top largest
// eigenpairs: values in top_vals[0..nev_top) (ascending, like LAPACK's tail) and vectors (m x nev_top, ldv) in top_vecs (device).
// want: bit0 = smallest eigenvalue must converge, bit1 = largest nev_top must converge.
// kmax_cap bounds the Krylov dimension (default 500; smaller values are a test hook for the non-convergence path).
LanczosResult lanczos_extreme(const double* T, int m, int ld, int want, int nev_top, double* top_vals_host, double* top_vecs,
                              int ldv, double tol, LanczosWork& w, cudaStream_t st, int kmax_cap = 500);

// Batched smallest eigenvalue of small symmetric matrices (64 < m <= 512 is the intended range, any m >= 1 works): one CTA
// per matrix reduces it to tridiagonal form by unblocked Householder reflections (matrix stays in L2, DESTROYED) and finds
// the smallest eigenvalue by parallel multisection on Sturm counts.  Replaces `eigmin` for multi-block problems where a
// Lanczos run per block would be launch-latency bound.  mats: device array of matrix pointers (full symmetric storage).
void batched_lambda_min(double* const* mats, const int* ms, const int* lds, int count, double* out, cudaStream_t st);

// Host-side symmetric tridiagonal eigen-solver (implicit QL). d[k] diag, e[k-1] offdiag. On return d = eigenvalues ascending;
// if Z != nullptr it must be k x k (row-major identity on input not required) and receives the eigenvectors as columns
// (Z[i*k + j] = component i of vector j); if zlast != nullptr it receives the last components of every eigenvector.
bool tridiag_ql(int k, double* d, double* e, double* Z, double* zlast);

}  // namespace lrn
