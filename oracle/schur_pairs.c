/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path (loraine.jl_b200/).
 *
 * Plain-C restatement of the sparse Schur-complement assembly of Loraine.jl for one PSD block,
 *     H[j,k] = tr(calA_j W calA_k W) = sum_{(a,b) in j} sum_{(p,q) in k} v_j[a,b] v_k[p,q] W[a,p] W[b,q],
 * i.e. the quantity the F3 branch of makeBBBBsi evaluates pair by pair through `_dot`
 * (/root/reference/src/makeBBBB.jl:139-213 and :39-64); every branch of that function (F1 :81-104, F3) computes this same
 * matrix.  Only the lower triangle (k <= j) is written, like `Hermitian(BBBB, :L)` reads it
 * (/root/reference/src/predictor_corrector.jl:39).  The NumPy oracle (oracle/loraine_oracle.py: makeBBBBsi_entries and the
 * literal makeBBBBsi_aswritten) pins this file in tests/test_oracle_golden.py; it exists so that the CPU arm of bench.py can
 * time the reference formulation at the full size of configs[4] (n_var = 40000: 8e8 pairs), where the NumPy
 * version needs minutes.  pthreads over rows j, dynamic schedule (row j costs nnz_j * sum_{k<=j} nnz_k).
 *
 * Layout: constraint-major entry lists (CSR of AA_i, math sign): entries rowptr[j] .. rowptr[j+1]-1 of constraint j are
 * (ep[e], eq[e], ev[e]) = (row, column, value) of calA_j, 0-based; W is m x m column-major (symmetric); H is n x n
 * column-major with leading dimension ldh.  accumulate != 0 adds to H (several PSD blocks), else overwrites.
 * [k0, k1) restricts the work to a column panel (k1 <= 0: all columns); column k is then stored at H + (k - k0) * ldh.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdatomic.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
    int64_t n, m, ldh;
    const int64_t* rowptr;
    const int32_t *ep, *eq;
    const double *ev, *W;
    double* H;
    int32_t accumulate;
    int64_t k0, k1;            /* columns [k0, k1) only; column k is stored at H + (k - k0) * ldh */
    atomic_long next;          /* dynamic schedule: chunks of 16 rows, long rows first */
} job_t;

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    const int64_t n = J->n, m = J->m;
    for (;;) {
        const int64_t c0 = atomic_fetch_add(&J->next, 16);
        if (c0 >= n) break;
        const int64_t c1 = c0 + 16 < n ? c0 + 16 : n;
        for (int64_t jj = c0; jj < c1; jj++) {
            const int64_t j = n - 1 - jj;
            const int64_t e0 = J->rowptr[j], e1 = J->rowptr[j + 1];
            if (e0 == e1) continue;
            const int64_t kend = j + 1 < J->k1 ? j + 1 : J->k1;
            for (int64_t k = J->k0; k < kend; k++) {
                const int64_t f0 = J->rowptr[k], f1 = J->rowptr[k + 1];
                if (f0 == f1) continue;
                double acc = 0.0;
                for (int64_t e = e0; e < e1; e++) {
                    const double* Wa = J->W + (size_t)J->ep[e] * m;   /* column a of the symmetric W: W[p, a] = Wa[p] */
                    const double* Wb = J->W + (size_t)J->eq[e] * m;
                    double s = 0.0;
                    for (int64_t f = f0; f < f1; f++) s += J->ev[f] * Wa[J->ep[f]] * Wb[J->eq[f]];
                    acc += J->ev[e] * s;
                }
                double* dst = J->H + (size_t)(k - J->k0) * J->ldh + j;
                if (J->accumulate) *dst += acc; else *dst = acc;
            }
        }
    }
    return NULL;
}

int32_t lrn_oracle_max_threads(void) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return (int32_t)(c < 1 ? 1 : (c > 256 ? 256 : c));
}

void lrn_oracle_schur_pairs(int64_t n, int64_t m, const int64_t* rowptr, const int32_t* ep, const int32_t* eq, const double* ev,
                            const double* W, double* H, int64_t ldh, int32_t accumulate, int32_t nthreads, int64_t k0, int64_t k1) {
    job_t J = {n, m, ldh, rowptr, ep, eq, ev, W, H, accumulate, k0 < 0 ? 0 : k0, (k1 <= 0 || k1 > n) ? n : k1};
    atomic_init(&J.next, 0);
    if (nthreads <= 0) nthreads = lrn_oracle_max_threads();
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 1; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, worker, &J) == 0) started++;
    worker(&J);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}
