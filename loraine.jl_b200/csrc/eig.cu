#include "eig.cuh"
#include "gemm.cuh"
#include <algorithm>
#include <cmath>

namespace lrn {
namespace {

constexpr int EN = 64, ELD = 65;
constexpr size_t EIG_SMEM = (2 * EN * ELD + 64) * sizeof(double) + 64 * sizeof(int);

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    // valid for non-negative finite doubles: bit patterns are ordered like unsigned integers
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// Batched two-sided cyclic Jacobi with round-robin parallel ordering, everything in shared memory.
__global__ void __launch_bounds__(256) jacobi_eig64_kernel(const EigSmallParams P) {
    extern __shared__ double sm[];
    double* a = sm;
    double* v = sm + EN * ELD;
    double* cs = v + EN * ELD;      // 32
    double* sn = cs + 32;           // 32
    int* pp = reinterpret_cast<int*>(sn + 32);
    int* qq = pp + 32;
    __shared__ int rotated;
    __shared__ double red[32];

    const int tid = threadIdx.x, z = blockIdx.x;
    const int n = P.n, n2 = (n + 1) & ~1, h = n2 >> 1;
    const double* Az = P.A + (size_t)z * P.sA;

    if (P.dg_in) {
        // diagonal blocks from the previous round's rotated Gram matrices, cross block from the split-K partial products
        const double* d0 = P.dg_in + (size_t)(2 * z) * 1024;
        const double* d1 = d0 + 1024;
        for (int idx = tid; idx < 1024; idx += 256) {
            int i = idx & 31, j = idx >> 5;
            a[i * ELD + j] = d0[idx];
            a[(32 + i) * ELD + 32 + j] = d1[idx];
            double val = 0.0;
            for (int t = 0; t < P.nparts; t++) val += Az[(size_t)t * P.sPart + (size_t)j * P.lda + i];
            a[(32 + i) * ELD + j] = val;
            a[j * ELD + 32 + i] = val;
        }
        for (int idx = tid; idx < EN * EN; idx += 256) {
            int i = idx % EN, j = idx / EN;
            v[i * ELD + j] = (i == j) ? 1.0 : 0.0;
        }
    } else {
        for (int idx = tid; idx < EN * EN; idx += 256) {
            int i = idx % EN, j = idx / EN;
            double val = 0.0;
            if (i < n && j < n)
                for (int t = 0; t < P.nparts; t++) val += Az[(size_t)t * P.sPart + (size_t)j * P.lda + i];
            a[i * ELD + j] = val;
            v[i * ELD + j] = (i == j) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < EN * EN; idx += 256) {
        int i = idx % EN, j = idx / EN;
        if (i > j) {
            double t = 0.5 * (a[i * ELD + j] + a[j * ELD + i]);
            a[i * ELD + j] = t;
            a[j * ELD + i] = t;
        }
    }
    __syncthreads();
    double lsum = 0.0, lmax = 0.0;
    for (int idx = tid; idx < EN * EN; idx += 256) {
        int i = idx % EN, j = idx / EN;
        double x = a[i * ELD + j];
        lsum += x * x;
        if (i < j && x != 0.0) {
            double den = sqrt(fabs(a[i * ELD + i] * a[j * ELD + j]));
            double r = (den > 0.0) ? fabs(x) / den : 1.0e300;
            lmax = fmax(lmax, r);
        }
    }
    const double fro = sqrt(block_sum(lsum, red));
    if (P.offmax) {
        lmax = warp_max(lmax);
        if ((tid & 31) == 0) atomic_max_nonneg(P.offmax, fmin(lmax, 1.0e300));
    }
    const double abs_tol = P.relative ? 0.0 : 1.0e-18 * fro;
    const double rel_tol = 1.0e-15;

    for (int sweep = 0; sweep < P.max_sweeps; sweep++) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        const int nsteps = P.cross_only ? 32 : n2 - 1;
        for (int r = 0; r < nsteps; r++) {
            // every warp owns 4 of the (at most 32) disjoint pairs of this step: lanes 0..3 compute the rotations of the
            // warp's own pairs (they only read rows p,q that no other warp writes in the row phase), no block barrier needed
            const int lane = tid & 31, wk0 = (tid >> 5) * 4;
            double c_l = 1.0, s_l = 0.0;
            int p_l = 0, q_l = 0;
            if (lane < 4 && wk0 + lane < h) {
                const int k = wk0 + lane;
                int p, q;
                if (P.cross_only) { p = k; q = 32 + ((k + r) & 31); }
                else if (k == 0) { p = n2 - 1; q = r; }
                else { p = (r + k) % (n2 - 1); q = (r - k + (n2 - 1)) % (n2 - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                if (q < n) {
                    double apq = a[p * ELD + q], app = a[p * ELD + p], aqq = a[q * ELD + q];
                    double thr = fmax(rel_tol * sqrt(fabs(app * aqq)), abs_tol);
                    if (fabs(apq) > thr) {
                        double theta = (aqq - app) / (2.0 * apq);
                        double t;
                        if (fabs(theta) > 1.0e100) t = 0.5 / theta;
                        else t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c_l = rsqrt(t * t + 1.0);
                        s_l = t * c_l;
                        rotated = 1;
                    }
                }
                p_l = p; q_l = q;
                cs[k] = c_l; sn[k] = s_l; pp[k] = p; qq[k] = q;       // for the column phase of the other warps
            }
            for (int kk = 0; kk < 4; kk++) {                         // rows p,q  <-  J^T A   (own pairs)
                const double s = __shfl_sync(0xffffffffu, s_l, kk);
                const double c = __shfl_sync(0xffffffffu, c_l, kk);
                const int p = __shfl_sync(0xffffffffu, p_l, kk), q = __shfl_sync(0xffffffffu, q_l, kk);
                if (wk0 + kk >= h || s == 0.0) continue;
                for (int j = lane; j < n2; j += 32) {
                    double ap = a[p * ELD + j], aq = a[q * ELD + j];
                    a[p * ELD + j] = c * ap - s * aq;
                    a[q * ELD + j] = s * ap + c * aq;
                }
            }
            __syncthreads();
            for (int kk = 0; kk < 4; kk++) {                         // cols p,q  <-  A J ;  V J ; pivot entries zeroed
                const int k = wk0 + kk;
                if (k >= h) break;
                const double s = sn[k];
                if (s == 0.0) continue;
                const double c = cs[k];
                const int p = pp[k], q = qq[k];
                for (int i = lane; i < n2; i += 32) {
                    double ap = a[i * ELD + p], aq = a[i * ELD + q];
                    double np_ = c * ap - s * aq, nq_ = s * ap + c * aq;
                    if (i == p) nq_ = 0.0;
                    if (i == q) np_ = 0.0;
                    a[i * ELD + p] = np_;
                    a[i * ELD + q] = nq_;
                    double vp = v[i * ELD + p], vq = v[i * ELD + q];
                    v[i * ELD + p] = c * vp - s * vq;
                    v[i * ELD + q] = s * vp + c * vq;
                }
            }
            __syncthreads();
        }
        const int any_rot = rotated;
        __syncthreads();           // everyone has read the flag before thread 0 may reset it
        if (!any_rot) break;
    }

    // output (optionally sorted descending by eigenvalue; ties broken by index)
    if (tid < n) {
        double li = a[tid * ELD + tid];
        int rank = tid;
        if (P.sort_desc) {
            rank = 0;
            for (int j = 0; j < n; j++) {
                double lj = a[j * ELD + j];
                if (lj > li || (lj == li && j < tid)) rank++;
            }
        }
        pp[tid] = rank;     // pairs finished: reuse as rank table (n <= 64 needs 64 ints: pp+qq are contiguous)
        if (P.evals) P.evals[(size_t)z * P.sE + rank] = li;
    }
    if (P.minval) {
        double li = (tid < n) ? a[tid * ELD + tid] : 1.0e300;
        li = warp_min(li);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = li;
        __syncthreads();
        if (tid == 0) {
            double m = red[0];
            for (int w = 1; w < 8; w++) m = fmin(m, red[w]);
            P.minval[z] = m;
        }
    }
    __syncthreads();
    if (P.dg_out) {
        double* o0 = P.dg_out + (size_t)P.slotmap[2 * z] * 1024;
        double* o1 = P.dg_out + (size_t)P.slotmap[2 * z + 1] * 1024;
        for (int idx = tid; idx < 1024; idx += 256) {
            int i = idx & 31, j = idx >> 5;
            o0[idx] = 0.5 * (a[i * ELD + j] + a[j * ELD + i]);
            o1[idx] = 0.5 * (a[(32 + i) * ELD + 32 + j] + a[(32 + j) * ELD + 32 + i]);
        }
    }
    if (P.V) {
        double* Vz = P.V + (size_t)z * P.sV;
        for (int idx = tid; idx < n * n; idx += 256) {
            int i = idx % n, j = idx / n;
            Vz[(size_t)pp[j] * P.ldv + i] = v[i * ELD + j];
        }
    }
}


// Cross-pair sweep of a 64 x 64 pair matrix (block-Jacobi rounds > 0): 32 steps, step r rotates the 32 disjoint index
// pairs (k, 32 + (k + r) mod 32).  1024 threads: thread (ki, kj) owns the 2 x 2 block {p_i, q_i} x {p_j, q_j}, which the
// two-sided update J' A J maps onto itself, so a step needs no intermediate barrier between the row and the column
// rotation: warp 0 computes the 32 rotations, one barrier, every thread rotates its blocks of A and V, one barrier.
__global__ void __launch_bounds__(1024) jacobi_cross64_kernel(const EigSmallParams P) {
    extern __shared__ double sm[];
    double* a = sm;
    double* v = sm + EN * ELD;
    double* cs = v + EN * ELD;      // 32
    double* sn = cs + 32;           // 32
    const int tid = threadIdx.x, z = blockIdx.x, lane = tid & 31;
    const double* Az = P.A + (size_t)z * P.sA;

    if (P.dg_in) {
        const double* d0 = P.dg_in + (size_t)(2 * z) * 1024;
        const double* d1 = d0 + 1024;
        const int i = tid & 31, j = tid >> 5;
        a[i * ELD + j] = d0[tid];
        a[(32 + i) * ELD + 32 + j] = d1[tid];
        double val = 0.0;
        for (int t = 0; t < P.nparts; t++) val += Az[(size_t)t * P.sPart + (size_t)j * P.lda + i];
        a[(32 + i) * ELD + j] = val;
        a[j * ELD + 32 + i] = val;
    } else {
        for (int idx = tid; idx < EN * EN; idx += 1024) {
            const int i = idx & 63, j = idx >> 6;
            if (i >= j) {
                double lo = 0.0, up = 0.0;
                for (int t = 0; t < P.nparts; t++) {
                    lo += Az[(size_t)t * P.sPart + (size_t)j * P.lda + i];
                    up += Az[(size_t)t * P.sPart + (size_t)i * P.lda + j];
                }
                const double val = 0.5 * (lo + up);
                a[i * ELD + j] = val;
                a[j * ELD + i] = val;
            }
        }
    }
    for (int idx = tid; idx < EN * EN; idx += 1024) {
        const int i = idx & 63, j = idx >> 6;
        v[i * ELD + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (P.offmax) {
        double lmax = 0.0;
        for (int idx = tid; idx < EN * EN; idx += 1024) {
            const int i = idx & 63, j = idx >> 6;
            const double x = a[i * ELD + j];
            if (i < j && x != 0.0) {
                const double den = sqrt(fabs(a[i * ELD + i] * a[j * ELD + j]));
                lmax = fmax(lmax, (den > 0.0) ? fabs(x) / den : 1.0e300);
            }
        }
        lmax = warp_max(lmax);
        if (lane == 0 && lmax > 0.0) atomic_max_nonneg(P.offmax, fmin(lmax, 1.0e300));
    }
    const int ki = tid >> 5, kj = lane;
    for (int r = 0; r < 32; r++) {
        if (tid < 32) {
            const int p = tid, q = 32 + ((tid + r) & 31);
            const double apq = a[p * ELD + q], app = a[p * ELD + p], aqq = a[q * ELD + q];
            double c = 1.0, s = 0.0;
            if (fabs(apq) > 1.0e-15 * sqrt(fabs(app * aqq))) {
                // t = sign(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq), without forming theta
                const double d = aqq - app, x = 2.0 * apq;
                const double h = sqrt(d * d + x * x);
                const double t = x / (d + copysign(h, d));
                c = rsqrt(t * t + 1.0);
                s = t * c;
            }
            cs[tid] = c;
            sn[tid] = s;
        }
        __syncthreads();
        {
            const int pi = ki, qi = 32 + ((ki + r) & 31), pj = kj, qj = 32 + ((kj + r) & 31);
            const double ci = cs[ki], si = sn[ki], cj = cs[kj], sj = sn[kj];
            const double x11 = a[pi * ELD + pj], x12 = a[pi * ELD + qj], x21 = a[qi * ELD + pj], x22 = a[qi * ELD + qj];
            const double y11 = ci * x11 - si * x21, y12 = ci * x12 - si * x22;
            const double y21 = si * x11 + ci * x21, y22 = si * x12 + ci * x22;
            double z11 = cj * y11 - sj * y12, z12 = sj * y11 + cj * y12;
            double z21 = cj * y21 - sj * y22, z22 = sj * y21 + cj * y22;
            if (ki == kj && si != 0.0) { z12 = 0.0; z21 = 0.0; }
            a[pi * ELD + pj] = z11; a[pi * ELD + qj] = z12; a[qi * ELD + pj] = z21; a[qi * ELD + qj] = z22;
            const double v11 = v[pi * ELD + pj], v12 = v[pi * ELD + qj], v21 = v[qi * ELD + pj], v22 = v[qi * ELD + qj];
            v[pi * ELD + pj] = cj * v11 - sj * v12; v[pi * ELD + qj] = sj * v11 + cj * v12;
            v[qi * ELD + pj] = cj * v21 - sj * v22; v[qi * ELD + qj] = sj * v21 + cj * v22;
        }
        __syncthreads();
    }
    if (P.dg_out) {
        double* o0 = P.dg_out + (size_t)P.slotmap[2 * z] * 1024;
        double* o1 = P.dg_out + (size_t)P.slotmap[2 * z + 1] * 1024;
        const int i = tid & 31, j = tid >> 5;
        o0[tid] = 0.5 * (a[i * ELD + j] + a[j * ELD + i]);
        o1[tid] = 0.5 * (a[(32 + i) * ELD + 32 + j] + a[(32 + j) * ELD + 32 + i]);
    }
    if (P.V) {
        double* Vz = P.V + (size_t)z * P.sV;
        for (int idx = tid; idx < EN * EN; idx += 1024) {
            const int i = idx & 63, j = idx >> 6;
            Vz[(size_t)j * P.ldv + i] = v[i * ELD + j];
        }
    }
}


// Register-resident variant of the cross-pair sweep for the recycled-Gram rounds (diagonal blocks from dg_in, cross block
// from the split-K partials).  Thread (ki = warp, kj = lane) keeps A11[ki][kj], the travelling entry A12[ki][(kj+r)%32] and
// its four entries of V in registers; the entries that move between warps (A22 and the transposed read of A12) go through
// two 32 x 33 shared arrays.  The column index a thread needs next step is the one its lane neighbour holds now, so the
// travelling values advance by one warp shuffle per step.  V never touches shared memory, which lets the V update of step
// r overlap with warp 0 computing the rotations of step r + 1 (double-buffered c / s).
__global__ void __launch_bounds__(1024) jacobi_cross64_reg_kernel(const EigSmallParams P) {
    // A12 / A22 / diag(A11) are double-buffered: step r reads buffer r & 1 and writes the other one (every entry is rewritten
    // in every step), so a step needs two barriers: rotations visible, new entries visible
    __shared__ double S12[2][32 * 33], S22[2][32 * 33], d1[2][32], cs[2][32], sn[2][32];
    extern __shared__ double vout[];                 // 64 x 65 staging of V for coalesced stores
    const int tid = threadIdx.x, z = blockIdx.x, lane = tid & 31, ki = tid >> 5, kj = lane;
    const double* Az = P.A + (size_t)z * P.sA;
    const double* g0 = P.dg_in + (size_t)(2 * z) * 1024;
    const double* g1 = g0 + 1024;
    double a11 = g0[tid];                            // A11[ki][kj] (symmetric block, read along the contiguous index)
    double a12 = 0.0;                                // A12[ki][kj] = <column ki of block 2z, column kj of block 2z+1>
    for (int t = 0; t < P.nparts; t++) a12 += Az[(size_t)t * P.sPart + (size_t)ki * P.lda + kj];
    S22[0][lane * 33 + ki] = g1[tid];                // A22[lane][ki] = g1[lane + 32 ki]
    S12[0][ki * 33 + kj] = a12;
    if (ki == kj) d1[0][ki] = a11;
    double v11 = (ki == kj) ? 1.0 : 0.0, v12 = 0.0, v21 = 0.0, v22 = v11;
    __syncthreads();
    if (P.offmax) {
        const double dii = d1[0][ki], djj = d1[0][kj], eii = S22[0][ki * 33 + ki], ejj = S22[0][kj * 33 + kj];
        const double a22 = S22[0][ki * 33 + kj];
        double lmax = 0.0, den;
        if (ki < kj && a11 != 0.0) { den = sqrt(fabs(dii * djj)); lmax = fmax(lmax, den > 0.0 ? fabs(a11) / den : 1.0e300); }
        if (a12 != 0.0) { den = sqrt(fabs(dii * ejj)); lmax = fmax(lmax, den > 0.0 ? fabs(a12) / den : 1.0e300); }
        if (ki < kj && a22 != 0.0) { den = sqrt(fabs(eii * ejj)); lmax = fmax(lmax, den > 0.0 ? fabs(a22) / den : 1.0e300); }
        lmax = warp_max(lmax);
        if (lane == 0 && lmax > 0.0) atomic_max_nonneg(P.offmax, fmin(lmax, 1.0e300));
    }
    const int nb1 = (lane + 1) & 31;
    double cjv = 1.0, sjv = 0.0;                      // column rotation of the previous step (V update is deferred)
    for (int r = 0; r <= 32; r++) {
        const int buf = r & 1;
        if (r < 32 && ki == 0) {
            const int q = (lane + r) & 31;
            const double apq = S12[buf][lane * 33 + q], app = d1[buf][lane], aqq = S22[buf][q * 33 + q];
            double c = 1.0, s = 0.0;
            if (fabs(apq) > 1.0e-15 * sqrt(fabs(app * aqq))) {
                const double d = aqq - app, x = 2.0 * apq;
                const double h = sqrt(d * d + x * x);
                const double t = x / (d + copysign(h, d));
                c = rsqrt(t * t + 1.0);
                s = t * c;
            }
            cs[buf][lane] = c;
            sn[buf][lane] = s;
        }
        if (r > 0) {
            // V (:, {kj, 32 + (kj + r - 1) % 32}) <- V J of the previous step, then the travelling columns move one lane on
            const double n11 = cjv * v11 - sjv * v12, n12 = sjv * v11 + cjv * v12;
            const double n21 = cjv * v21 - sjv * v22, n22 = sjv * v21 + cjv * v22;
            v11 = n11; v21 = n21;
            v12 = __shfl_sync(0xffffffffu, n12, nb1);
            v22 = __shfl_sync(0xffffffffu, n22, nb1);
        }
        if (r == 32) break;
        __syncthreads();
        {
            const int qi = (ki + r) & 31, qj = (kj + r) & 31;
            const double ci = cs[buf][ki], si = sn[buf][ki], cj = cs[buf][kj], sj = sn[buf][kj];
            const double x11 = a11, x12 = a12, x21 = S12[buf][kj * 33 + qi], x22 = S22[buf][qi * 33 + qj];
            const double y11 = ci * x11 - si * x21, y12 = ci * x12 - si * x22;
            const double y21 = si * x11 + ci * x21, y22 = si * x12 + ci * x22;
            const double z11 = cj * y11 - sj * y12;
            double z12 = sj * y11 + cj * y12;
            const double z22 = sj * y21 + cj * y22;
            if (ki == kj && si != 0.0) z12 = 0.0;
            a11 = z11;
            cjv = cj; sjv = sj;
            S12[buf ^ 1][ki * 33 + qj] = z12;
            S22[buf ^ 1][qi * 33 + qj] = z22;
            if (ki == kj) d1[buf ^ 1][ki] = z11;
            a12 = __shfl_sync(0xffffffffu, z12, nb1);
        }
        __syncthreads();
    }
    // outputs: rotated diagonal blocks for the next round, V (staged through shared memory for coalesced stores)
    double* o0 = P.dg_out + (size_t)P.slotmap[2 * z] * 1024;
    double* o1 = P.dg_out + (size_t)P.slotmap[2 * z + 1] * 1024;
    o0[tid] = a11;
    o1[tid] = 0.5 * (S22[0][lane * 33 + ki] + S22[0][ki * 33 + lane]);
    vout[ki * ELD + kj] = v11;
    vout[(32 + ki) * ELD + kj] = v21;
    vout[ki * ELD + 32 + kj] = v12;
    vout[(32 + ki) * ELD + 32 + kj] = v22;
    __syncthreads();
    double* Vz = P.V + (size_t)z * P.sV;
    for (int idx = tid; idx < EN * EN; idx += 1024) {
        const int i = idx & 63, j = idx >> 6;
        Vz[(size_t)j * P.ldv + i] = vout[i * ELD + j];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// one-sided block Jacobi helpers
// ---------------------------------------------------------------------------------------------------------------
__global__ void svd_init_kernel(const double* __restrict__ A, int lda, int m, int mp, int rows, double* __restrict__ W, int ldw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;    // row in stacked buffer
    int j = blockIdx.y;                               // column
    if (i >= rows) return;
    double v;
    if (i < m) v = (j < m) ? A[(size_t)j * lda + i] : 0.0;
    else v = (i - m == j) ? 1.0 : 0.0;
    W[(size_t)j * ldw + i] = v;
}

// max_{i>j} |C_ij| / sqrt(C_ii C_jj) of the lower triangle of a Gram matrix (exact orthogonality measure of the columns)
__global__ void __launch_bounds__(256) gram_offmax_kernel(const double* __restrict__ Cm, int ldc, int m, double* __restrict__ out) {
    const int j = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    double r = 0.0;
    if (i < m && i > j) {
        const double x = Cm[(size_t)j * ldc + i];
        if (x != 0.0) {
            const double den = sqrt(fabs(Cm[(size_t)i * ldc + i] * Cm[(size_t)j * ldc + j]));
            r = (den > 0.0) ? fabs(x) / den : 1.0e300;
        }
    }
    r = warp_max(r);
    if ((threadIdx.x & 31) == 0 && r > 0.0) atomic_max_nonneg(out, fmin(r, 1.0e300));
}

__global__ void __launch_bounds__(256) colnorm_kernel(const double* __restrict__ W, int ldw, int m, int ncols, double* __restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= ncols) return;
    const double* c = W + (size_t)warp * ldw;
    // scaled two-pass not needed: entries are O(sigma), squares stay in range for IPM data
    double s = 0.0;
    for (int i = lane; i < m; i += 32) s += c[i] * c[i];
    s = warp_sum(s);
    if (lane == 0) out[warp] = sqrt(s);
}

__global__ void rank_desc_kernel(const double* __restrict__ v, int n, int* __restrict__ perm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double vi = v[i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
        double vj = v[j];
        if (vj > vi || (vj == vi && j < i)) rank++;
    }
    perm[rank] = i;
}

__global__ void svd_gather_kernel(const double* __restrict__ W, int ldw, int m, const int* __restrict__ perm,
                                  const double* __restrict__ sv, double* __restrict__ UD, int ldu, double* __restrict__ V,
                                  int ldv, double* __restrict__ sigma) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    int src = perm[r];
    if (i < m) {
        UD[(size_t)r * ldu + i] = W[(size_t)src * ldw + i];
        if (V) V[(size_t)r * ldv + i] = W[(size_t)src * ldw + m + i];
    }
    if (i == 0) sigma[r] = sv[src];
}

__global__ void svd_init_batched_kernel(const double* const* __restrict__ A, int lda, int m, int mp, double* __restrict__ W, int ldw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, b = blockIdx.z;
    if (i >= m) return;
    W[((size_t)b * mp + j) * ldw + i] = (j < m) ? A[b][(size_t)j * lda + i] : 0.0;
}
__global__ void rank_desc_batched_kernel(const double* __restrict__ v, int n, int* __restrict__ perm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= n) return;
    const double* vb = v + (size_t)b * n;
    double vi = vb[i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
        double vj = vb[j];
        if (vj > vi || (vj == vi && j < i)) rank++;
    }
    perm[(size_t)b * n + rank] = i;
}
__global__ void svd_gather_batched_kernel(const double* __restrict__ W, int ldw, int m, int mp, const int* __restrict__ perm,
                                          const double* __restrict__ sv, double* const* __restrict__ UD, int ldu,
                                          double* const* __restrict__ sigma) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, b = blockIdx.z;
    int src = perm[(size_t)b * mp + r];
    if (i < m) UD[b][(size_t)r * ldu + i] = W[((size_t)b * mp + src) * ldw + i];
    if (i == 0) sigma[b][r] = sv[(size_t)b * mp + src];
}

// ---------------------------------------------------------------------------------------------------------------
// Lanczos kernels
// ---------------------------------------------------------------------------------------------------------------
// y[c] = sum_i A[i + c*ld] * x[i]   (one warp per column; for symmetric A this is A*x with fully coalesced reads)
__global__ void __launch_bounds__(256) gemv_t_kernel(const double* __restrict__ A, int ld, int rows, int cols,
                                                     const double* __restrict__ x, double* __restrict__ y) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= cols) return;
    const double* c = A + (size_t)warp * ld;
    double s = 0.0;
    if ((ld & 1) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
        // 16-byte loads, four independent partial sums: 2 KB of the column in flight per warp
        const double2* c2 = reinterpret_cast<const double2*>(c);
        const double2* x2 = reinterpret_cast<const double2*>(x);
        const int n2 = rows >> 1;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int i = lane;
        for (; i + 96 < n2; i += 128) {
            const double2 a0 = c2[i], a1 = c2[i + 32], a2 = c2[i + 64], a3 = c2[i + 96];
            const double2 b0 = x2[i], b1 = x2[i + 32], b2 = x2[i + 64], b3 = x2[i + 96];
            s0 += a0.x * b0.x + a0.y * b0.y;
            s1 += a1.x * b1.x + a1.y * b1.y;
            s2 += a2.x * b2.x + a2.y * b2.y;
            s3 += a3.x * b3.x + a3.y * b3.y;
        }
        for (; i < n2; i += 32) {
            const double2 a0 = c2[i], b0 = x2[i];
            s0 += a0.x * b0.x + a0.y * b0.y;
        }
        s = (s0 + s1) + (s2 + s3);
        if ((rows & 1) && lane == 0) s += c[rows - 1] * x[rows - 1];
    } else {
        for (int i = lane; i < rows; i += 32) s += c[i] * x[i];
    }
    s = warp_sum(s);
    if (lane == 0) y[warp] = s;
}
// Skinny variant for the re-orthogonalisation (few columns, long rows): grid (cols, RSPLIT), every CTA reduces one row chunk
// of one column; the partial sums are added up by the consumer (deterministic, no atomics).
constexpr int RSPLIT = 8;
__global__ void __launch_bounds__(128) gemv_t_split_kernel(const double* __restrict__ A, int ld, int rows,
                                                           const double* __restrict__ x, double* __restrict__ part) {
    __shared__ double red[32];
    const int col = blockIdx.x, sp = blockIdx.y;
    const int chunk = ((rows + RSPLIT - 1) / RSPLIT + 1) & ~1;
    const int r0 = sp * chunk, r1 = min(rows, r0 + chunk);
    const double* c = A + (size_t)col * ld;
    double s0 = 0.0, s1 = 0.0;
    int i = r0 + threadIdx.x;
    for (; i + 128 < r1; i += 256) { s0 += c[i] * x[i]; s1 += c[i + 128] * x[i + 128]; }
    if (i < r1) s0 += c[i] * x[i];
    const double s = block_sum(s0 + s1, red);
    if (threadIdx.x == 0) part[col * RSPLIT + sp] = s;
}
// w[i] -= sum_k Q[i + k*ld] * (sum of the RSPLIT partials of c[k])
__global__ void __launch_bounds__(128) gemv_n_sub_split_kernel(const double* __restrict__ Q, int ld, int rows, int cols,
                                                               const double* __restrict__ part, double* __restrict__ w) {
    extern __shared__ double cs_[];
    for (int k = threadIdx.x; k < cols; k += 128) {
        double t = 0.0;
#pragma unroll
        for (int u = 0; u < RSPLIT; u++) t += part[k * RSPLIT + u];
        cs_[k] = t;
    }
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= rows) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = 0;
    for (; k + 3 < cols; k += 4) {
        s0 += Q[(size_t)k * ld + i] * cs_[k];
        s1 += Q[(size_t)(k + 1) * ld + i] * cs_[k + 1];
        s2 += Q[(size_t)(k + 2) * ld + i] * cs_[k + 2];
        s3 += Q[(size_t)(k + 3) * ld + i] * cs_[k + 3];
    }
    for (; k < cols; k++) s0 += Q[(size_t)k * ld + i] * cs_[k];
    w[i] -= (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(1024) lanczos_init_kernel(double* __restrict__ q, int m) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < m; i += 1024) {
        unsigned hsh = (unsigned)i * 2654435761u + 12345u;
        hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
        double x = 0.5 + (double)(hsh & 0xffffu) / 65536.0;     // in [0.5, 1.5): never orthogonal to a Perron-like vector
        if (hsh & 0x10000u) x = -x;
        q[i] = x;
        s += x * x;
    }
    s = block_sum(s, red);
    double inv = rsqrt(s);
    for (int i = threadIdx.x; i < m; i += 1024) q[i] *= inv;
}
// alpha = q_j . w ; w -= alpha*q_j + beta_prev*q_{j-1}
__global__ void __launch_bounds__(1024) lanczos_alpha_kernel(double* __restrict__ w, const double* __restrict__ qj,
                                                             const double* __restrict__ qjm1,
                                                             const double* __restrict__ beta_prev_p, int m,
                                                             double* __restrict__ scal) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < m; i += 1024) s += qj[i] * w[i];
    double alpha = block_sum(s, red);
    const double beta_prev = qjm1 ? *beta_prev_p : 0.0;
    for (int i = threadIdx.x; i < m; i += 1024) {
        double x = w[i] - alpha * qj[i];
        if (qjm1) x -= beta_prev * qjm1[i];
        w[i] = x;
    }
    if (threadIdx.x == 0) scal[0] = alpha;
}
// beta = ||w|| ; q_next = w / beta
__global__ void __launch_bounds__(1024) lanczos_beta_kernel(const double* __restrict__ w, double* __restrict__ qn, int m,
                                                            double* __restrict__ scal) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < m; i += 1024) s += w[i] * w[i];
    double beta = sqrt(block_sum(s, red));
    double inv = beta > 0.0 ? 1.0 / beta : 0.0;
    for (int i = threadIdx.x; i < m; i += 1024) qn[i] = w[i] * inv;
    if (threadIdx.x == 0) scal[0] = beta;
}

// ---------------------------------------------------------------------------------------------------------------
// batched Householder tridiagonalisation + Sturm multisection (one CTA per matrix)
// ---------------------------------------------------------------------------------------------------------------
constexpr int TRD_T = 512;       // threads
constexpr int TRD_MAXM = 1024;   // v, p vectors in shared memory

__global__ void __launch_bounds__(TRD_T) k_batched_lambda_min(double* const* __restrict__ mats, const int* __restrict__ ms,
                                                               const int* __restrict__ lds, double* __restrict__ out) {
    __shared__ double v[TRD_MAXM], pv[TRD_MAXM], d[TRD_MAXM], e[TRD_MAXM];
    __shared__ double red[32];
    __shared__ double s_alpha, s_beta, s_lo, s_hi;
    __shared__ int cnt[TRD_T + 1];
    const int z = blockIdx.x, tid = threadIdx.x;
    double* A = mats[z];
    const int m = ms[z], ld = lds[z];
    for (int k = 0; k < m; k++) {
        const int r = m - k - 1;                       // length of the column below the diagonal
        if (tid == 0) d[k] = A[(size_t)k * ld + k];
        if (r <= 0) break;
        double* x = A + (size_t)k * ld + k + 1;        // column k below the diagonal
        double* A22 = A + (size_t)(k + 1) * ld + k + 1;
        if (r == 1) {
            if (tid == 0) e[k] = x[0];
            __syncthreads();
            continue;
        }
        // ||x||
        double s = 0.0;
        for (int i = tid; i < r; i += TRD_T) { double t = x[i]; v[i] = t; s += t * t; }
        s = block_sum(s, red);
        const double nrm = sqrt(s);
        if (nrm == 0.0) {
            if (tid == 0) e[k] = 0.0;
            __syncthreads();
            continue;
        }
        if (tid == 0) {
            const double x0 = v[0];
            const double alpha = (x0 >= 0.0) ? -nrm : nrm;
            v[0] = x0 - alpha;
            s_alpha = alpha;
            s_beta = 2.0 / (s - x0 * x0 + v[0] * v[0]);
            e[k] = alpha;
        }
        __syncthreads();
        const double beta = s_beta;
        // p = beta * A22 * v   (A22 symmetric, full storage: coalesced over rows)
        for (int i = tid; i < r; i += TRD_T) {
            double acc = 0.0;
            const double* row = A22 + i;
            for (int j = 0; j < r; j++) acc += row[(size_t)j * ld] * v[j];
            pv[i] = beta * acc;
        }
        __syncthreads();
        double pk = 0.0;
        for (int i = tid; i < r; i += TRD_T) pk += pv[i] * v[i];
        pk = block_sum(pk, red);
        const double K = 0.5 * beta * pk;
        for (int i = tid; i < r; i += TRD_T) pv[i] -= K * v[i];        // w = p - K v
        __syncthreads();
        // A22 <- A22 - v w' - w v'
        for (int idx = tid; idx < r * r; idx += TRD_T) {
            const int i = idx % r, j = idx / r;
            A22[(size_t)j * ld + i] -= v[i] * pv[j] + pv[i] * v[j];
        }
        __syncthreads();
    }
    __syncthreads();
    // Gershgorin bounds
    double lo = 1.0e300, hi = -1.0e300;
    for (int i = tid; i < m; i += TRD_T) {
        double rad = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < m - 1 ? fabs(e[i]) : 0.0);
        lo = fmin(lo, d[i] - rad);
        hi = fmax(hi, d[i] + rad);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = lo; }
    __syncthreads();
    if (tid == 0) { double t = red[0]; for (int w = 1; w < TRD_T / 32; w++) t = fmin(t, red[w]); s_lo = t; }
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = hi; }
    __syncthreads();
    if (tid == 0) { double t = red[0]; for (int w = 1; w < TRD_T / 32; w++) t = fmax(t, red[w]); s_hi = t; }
    __syncthreads();
    // multisection: thread t counts eigenvalues below x_t = lo + (t+1) (hi-lo)/(T+1); the smallest eigenvalue lies in the
    // first sub-interval whose right end has count >= 1
    for (int round = 0; round < 12; round++) {
        const double a = s_lo, b = s_hi;
        const double h = (b - a) / (TRD_T + 1);
        const double xs = a + (tid + 1) * h;
        int c = 0;
        double q = d[0] - xs;
        if (q < 0.0) c++;
        for (int i = 1; i < m; i++) {
            if (fabs(q) < 1.0e-300) q = -1.0e-300;
            q = d[i] - xs - e[i - 1] * e[i - 1] / q;
            if (q < 0.0) c++;
        }
        cnt[tid] = c;
        __syncthreads();
        if (tid == 0) {
            int first = TRD_T;                          // index of the first shift with count >= 1
            for (int t = 0; t < TRD_T; t++) if (cnt[t] >= 1) { first = t; break; }
            s_lo = a + first * h;                       // left neighbour (shift index first-1 -> a + first*h)
            s_hi = (first < TRD_T) ? a + (first + 1) * h : b;
        }
        __syncthreads();
        if (!(s_hi - s_lo > 4.0e-16 * fmax(fabs(s_lo), fabs(s_hi)))) break;
    }
    if (tid == 0) out[z] = 0.5 * (s_lo + s_hi);
}

}  // namespace

void batched_lambda_min(double* const* mats, const int* ms, const int* lds, int count, double* out, cudaStream_t st) {
    if (count <= 0) return;
    k_batched_lambda_min<<<count, TRD_T, 0, st>>>(mats, ms, lds, out);
    LRN_CHECK_LAUNCH();
}

void jacobi_eig_small(const EigSmallParams& p, cudaStream_t st) {
    LRN_REQUIRE(p.n >= 1 && p.n <= EN, "jacobi_eig_small handles n <= 64");
    static PerDeviceOnce once;               // the shared-memory opt-ins are per-device attributes
    once.run([&] {
        LRN_CUDA(cudaFuncSetAttribute(jacobi_eig64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EIG_SMEM));
        LRN_CUDA(cudaFuncSetAttribute(jacobi_cross64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EIG_SMEM));
        LRN_CUDA(cudaFuncSetAttribute(jacobi_cross64_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(EN * ELD * sizeof(double))));
    });
    if (p.cross_only && p.n == EN && p.relative && p.max_sweeps == 1 && !p.evals && !p.minval && !p.sort_desc) {
        if (p.dg_in && p.dg_out && p.V) {
            jacobi_cross64_reg_kernel<<<p.batch, 1024, EN * ELD * sizeof(double), st>>>(p);
            LRN_CHECK_LAUNCH();
            return;
        }
        jacobi_cross64_kernel<<<p.batch, 1024, EIG_SMEM, st>>>(p);
        LRN_CHECK_LAUNCH();
        return;
    }
    jacobi_eig64_kernel<<<p.batch, 256, EIG_SMEM, st>>>(p);
    LRN_CHECK_LAUNCH();
}

void SweepGraphs::reset() {
    for (auto& e : exec) { if (e) cudaGraphExecDestroy(e); e = nullptr; }
    nodes[0] = nodes[1] = 0;
    warm = false; broken = false;
}

// Runs one sweep: eagerly the first time on a workspace, then captured (thread-local capture: other host threads may drive other
// devices meanwhile) and replayed.  `body` enqueues the sweep on `st` and leaves the host-side ping-pong pointers advanced;
// `advance` advances them without enqueueing anything (after a replay).
template <typename Body, typename Advance>
static void run_sweep(SweepGraphs& G, int parity, cudaStream_t st, Body&& body, Advance&& advance) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread) cudaStreamIsCapturing(st, &cap);
    const bool graphs_ok = !G.broken && !gemm_profile_active() && cap == cudaStreamCaptureStatusNone;
    if (graphs_ok && G.exec[parity]) {
        LRN_CUDA(cudaGraphLaunch(G.exec[parity], st));
        g_kernel_launches.fetch_add(G.nodes[parity]);
        advance();
        return;
    }
    if (!graphs_ok || !G.warm) {
        body();
        G.warm = true;
        return;
    }
    // (the legacy default stream cannot be captured: callers that run on it, like the debug hooks, stay eager)
    bool ok = st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread &&
              cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        G.broken = true;
        body();
        return;
    }
    cudaGraph_t graph = nullptr;
    const long long before = g_kernel_launches.load();
    try { body(); } catch (...) { ok = false; }
    if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) ok = false;
    G.nodes[parity] = g_kernel_launches.load() - before;
    g_kernel_launches.fetch_sub(G.nodes[parity]);            // counted per replay
    if (ok && cudaGraphInstantiate(&G.exec[parity], graph, 0) != cudaSuccess) { ok = false; G.exec[parity] = nullptr; }
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        // nothing was enqueued by the captured body, but it advanced the host-side ping-pong pointers: put them back (the
        // advance is its own inverse) and run the sweep eagerly; stay eager from now on
        cudaGetLastError();
        G.broken = true;
        G.nodes[parity] = 0;
        advance();
        body();
        return;
    }
    LRN_CUDA(cudaGraphLaunch(G.exec[parity], st));
    g_kernel_launches.fetch_add(G.nodes[parity]);
}

void SvdWork::ensure(int m_, bool want_V_) {
    if (m_ == m && buf0.p && want_V_ == want_V) return;
    graphs.reset();
    m = m_;
    want_V = want_V_;
    mp = round_up(m, 64);
    const int rows_ = want_V ? m + mp : m;
    panel = rows_ >= 1024;                         // persistent TMA panel-rotation kernel: rows padded to whole 128-row tiles
    ldw = panel ? round_up(rows_, 128) + 8 : pad_ld(rows_);
    size_t elems = (size_t)ldw * mp;
    buf0.alloc(elems);
    buf1.alloc(elems);
    const int nblk = mp / 32, pairs = nblk / 2;
    splits = (int)std::min<long long>(16, std::max<long long>(1, cdiv(296, pairs)));   // >= 2 CTAs per SM in the Gram GEMM
    Kc = round_up((int)cdiv(m, splits), 16);
    if (Kc < 64) Kc = 64;
    splits = (int)cdiv(m, Kc);
    if (splits == 1) Kc = m;
    xsplits = (int)std::min<long long>(16, std::max<long long>(1, cdiv(444, pairs)));  // 3 CTAs per SM in the cross Gram GEMM
    xKc = round_up((int)cdiv(m, xsplits), 32);
    if (xKc < 256) xKc = 256;
    xsplits = (int)cdiv(m, xKc);
    gram.alloc((size_t)pairs * std::max(splits * 4096, xsplits * 1024));
    rot.alloc((size_t)pairs * 64 * 64);
    dg0.alloc((size_t)nblk * 1024);
    dg1.alloc((size_t)nblk * 1024);
    offmax.alloc(1);
    sv.alloc(mp);
    perm.alloc(mp);
    // round-robin chair rotation: pair k = slots (2k, 2k+1) = (t_k, b_k); t_0 fixed,
    // t_1 -> t_2 -> ... -> t_{h-1} -> b_{h-1} -> b_{h-2} -> ... -> b_0 -> t_1
    std::vector<int> pi(nblk);
    const int h = pairs;
    for (int s = 0; s < nblk; s++) pi[s] = s;
    if (h > 1) {
        pi[0] = 0;
        for (int k = 1; k < h - 1; k++) pi[2 * k] = 2 * (k + 1);
        pi[2 * (h - 1)] = 2 * (h - 1) + 1;
        for (int k = 1; k < h; k++) pi[2 * k + 1] = 2 * (k - 1) + 1;
        pi[1] = 2;
    }
    slotmap.upload(pi);
    inner_sweeps = (nblk == 2) ? 40 : 1;   // a single pair is diagonalised completely in one visit
    LRN_CUDA(cudaDeviceSynchronize());
}

int svd_block_jacobi(const double* A, int lda, int m, double* U_D, int ldu, double* V, int ldv, double* sigma, SvdWork& w,
                     double tol, int max_sweeps, cudaStream_t st) {
    const bool want_V = (V != nullptr);
    w.ensure(m, want_V);
    const int mp = w.mp, ldw = w.ldw, nblk = mp / 32, pairs = nblk / 2, rounds = nblk - 1;
    const int rows = want_V ? m + mp : m;
    static const bool trace = getenv("LRN_SVD_TRACE") != nullptr;
    double* cur = w.buf0.p;
    double* nxt = w.buf1.p;
    {
        dim3 grid((unsigned)cdiv(rows, 256), (unsigned)mp);
        svd_init_kernel<<<grid, 256, 0, st>>>(A, lda, m, mp, rows, cur, ldw);
        LRN_CHECK_LAUNCH();
    }
    const int splits = w.splits, Kc = w.Kc, xsplits = w.xsplits, xKc = w.xKc;
    const bool recycle = nblk > 2 && m >= 256;   // hand the diagonal Gram blocks from round to round (see EigSmallParams)
    double* dgc = w.dg0.p;
    double* dgn = w.dg1.p;
    int sweeps = 0;
    // the operand pointers of a sweep depend only on its parity (odd number of rounds: the ping-pong buffers end swapped)
    auto sweep_body = [&]() {
        LRN_CUDA(cudaMemsetAsync(w.offmax.p, 0, sizeof(double), st));
        for (int r = 0; r < rounds; r++) {
            // Gram matrices of all column-block pairs (split-K partials).  The first round of a sweep forms the full 64 x 64
            // products from the columns (this also bounds the drift of the recycled diagonal blocks to one sweep); the other
            // rounds only need the 32 x 32 cross block B_{2z+1}' B_{2z}
            const bool cross = recycle && r > 0;
            GemmParams g;
            g.A = cross ? cur + (size_t)32 * ldw : cur; g.B = cur; g.C = w.gram.p;
            g.transA = true; g.transB = false;
            g.lda = ldw; g.ldb = ldw;
            g.batch = pairs; g.sA = (long long)64 * ldw; g.sB = (long long)64 * ldw;
            if (cross) {
                g.M = 32; g.N = 32; g.K = xKc; g.ldc = 32; g.sC = (long long)xsplits * 1024;
                g.batch2 = xsplits; g.sA2 = xKc; g.sB2 = xKc; g.sC2 = 1024;
                g.K_last = m - (xsplits - 1) * xKc;
            } else {
                g.M = 64; g.N = 64; g.K = Kc; g.ldc = 64; g.sC = (long long)splits * 4096;
                g.batch2 = splits; g.sA2 = Kc; g.sB2 = Kc; g.sC2 = 4096;
                g.K_last = m - (splits - 1) * Kc;
            }
            gemm(g, st);
            EigSmallParams e;
            e.A = w.gram.p; e.n = 64;
            if (cross) { e.lda = 32; e.sA = (long long)xsplits * 1024; e.nparts = xsplits; e.sPart = 1024; e.dg_in = dgc; }
            else { e.lda = 64; e.sA = (long long)splits * 4096; e.nparts = splits; e.sPart = 4096; }
            if (recycle) { e.dg_out = dgn; e.slotmap = w.slotmap.p; std::swap(dgc, dgn); }
            e.V = w.rot.p; e.ldv = 64; e.sV = 4096; e.relative = 1; e.offmax = w.offmax.p; e.batch = pairs;
            e.max_sweeps = w.inner_sweeps;   // one cyclic sweep per visit converges in as many outer sweeps as full diagonalisation
            e.cross_only = (nblk > 2 && r > 0) ? 1 : 0;   // pairs inside a 32-column block: once per sweep (round 0) is enough
            // note: no sorting inside the pair rotations -- with the round-robin block ordering it makes columns migrate
            // between blocks and the sweep no longer visits every column pair (observed: no convergence)
            jacobi_eig_small(e, st);
            if (w.panel) {                      // rotate [A;V] panels and scatter them to next round's arrangement
                PanelRotateParams u;
                u.cur = cur; u.nxt = nxt; u.rot = w.rot.p; u.slotmap = w.slotmap.p; u.ldw = ldw;
                u.tiles = (int)cdiv(rows, 128); u.total = (long long)u.tiles * pairs;
                panel_rotate(u, st);
            } else {
                GemmParams u;
                u.A = cur; u.B = w.rot.p; u.C = nxt;
                u.M = rows; u.N = 64; u.K = 64; u.lda = ldw; u.ldb = 64; u.ldc = ldw;
                u.batch = pairs; u.sA = (long long)64 * ldw; u.sB = 4096;
                u.cblkmap = w.slotmap.p;
                gemm(u, st);
            }
            std::swap(cur, nxt);
        }
    };
    auto sweep_advance = [&]() {
        if (rounds & 1) { std::swap(cur, nxt); if (recycle) std::swap(dgc, dgn); }
    };
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
        run_sweep(w.graphs, (rounds & 1) ? (sweep & 1) : 0, st, sweep_body, sweep_advance);
        sweeps++;
        double off = 0.0;
        LRN_CUDA(cudaMemcpyAsync(&off, w.offmax.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        LRN_CUDA(cudaStreamSynchronize(st));
        if (trace) fprintf(stderr, "[lrn svd] m=%d sweep %d offmax %.3e\n", m, sweep + 1, off);
        if (off <= tol) break;
        // `off <= tol` at the start of a sweep leaves the columns orthogonal to ~1e2 tol^2 after it; the exact check below
        // accepts the same level (not tol itself), so both exits deliver the accuracy the Schur parity bound (1e-11) needs
        const double state_tol = 1.0e4 * tol * tol;
        if (w.panel && 1.0e2 * off * off <= state_tol) {
            // quadratic regime (measured: a sweep takes the measure from e to ~1e2 e^2): the sweep just finished has most
            // likely converged.  Measure the state exactly with one Gram product (a third of a sweep's time) instead of
            // spending a whole sweep on finding out: C = cur' cur into the idle buffer
            GemmParams c;
            c.A = cur; c.B = cur; c.C = nxt; c.transA = true; c.transB = false;
            c.M = m; c.N = m; c.K = m; c.lda = ldw; c.ldb = ldw; c.ldc = ldw; c.lower = 1;
            gemm(c, st);
            LRN_CUDA(cudaMemsetAsync(w.offmax.p, 0, sizeof(double), st));
            gram_offmax_kernel<<<dim3((unsigned)cdiv(m, 256), (unsigned)m), 256, 0, st>>>(nxt, ldw, m, w.offmax.p);
            LRN_CHECK_LAUNCH();
            double state = 0.0;
            LRN_CUDA(cudaMemcpyAsync(&state, w.offmax.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            LRN_CUDA(cudaStreamSynchronize(st));
            if (trace) fprintf(stderr, "[lrn svd] m=%d state after sweep %d: %.3e\n", m, sweep + 1, state);
            if (state <= state_tol) break;
        }
    }
    // after a whole number of sweeps the arrangement is back to the identity; singular values = column norms
    colnorm_kernel<<<(unsigned)cdiv((long long)mp * 32, 256), 256, 0, st>>>(cur, ldw, m, mp, w.sv.p);
    LRN_CHECK_LAUNCH();
    rank_desc_kernel<<<(unsigned)cdiv(mp, 256), 256, 0, st>>>(w.sv.p, mp, w.perm.p);
    LRN_CHECK_LAUNCH();
    dim3 grid((unsigned)cdiv(m, 256), (unsigned)m);
    svd_gather_kernel<<<grid, 256, 0, st>>>(cur, ldw, m, w.perm.p, w.sv.p, U_D, ldu, V, ldv, sigma);
    LRN_CHECK_LAUNCH();
    return sweeps;
}

void SvdBatchWork::ensure(int m_, int nb_) {
    if (m_ == m && nb_ == nb && buf0.p) return;
    graphs.reset();
    m = m_; nb = nb_;
    mp = round_up(m, 64);
    ldw = pad_ld(m);
    buf0.alloc((size_t)ldw * mp * nb);
    buf1.alloc((size_t)ldw * mp * nb);
    const int nblk = mp / 32, pairs = nblk / 2;
    splits = (int)std::min<long long>(16, std::max<long long>(1, cdiv(296, (long long)pairs * nb)));
    Kc = round_up((int)cdiv(m, splits), 16);
    if (Kc < 64) Kc = 64;
    splits = (int)cdiv(m, Kc);
    if (splits == 1) Kc = m;
    gram.alloc((size_t)pairs * nb * splits * 64 * 64);
    rot.alloc((size_t)pairs * nb * 64 * 64);
    offmax.alloc(1);
    sv.alloc((size_t)mp * nb);
    perm.alloc((size_t)mp * nb);
    std::vector<int> pi(nblk), all((size_t)nblk * nb);
    const int h = pairs;
    for (int s = 0; s < nblk; s++) pi[s] = s;
    if (h > 1) {
        pi[0] = 0;
        for (int k = 1; k < h - 1; k++) pi[2 * k] = 2 * (k + 1);
        pi[2 * (h - 1)] = 2 * (h - 1) + 1;
        for (int k = 1; k < h; k++) pi[2 * k + 1] = 2 * (k - 1) + 1;
        pi[1] = 2;
    }
    for (int b = 0; b < nb; b++)
        for (int s = 0; s < nblk; s++) all[(size_t)b * nblk + s] = b * nblk + pi[s];
    slotmap.upload(all);
    LRN_CUDA(cudaDeviceSynchronize());
}

int svd_block_jacobi_batched(const double* const* A, int lda, int m, int nb, double* const* UD, int ldu, double* const* sigma,
                             SvdBatchWork& w, double tol, int max_sweeps, cudaStream_t st) {
    w.ensure(m, nb);
    const int mp = w.mp, ldw = w.ldw, nblk = mp / 32, pairs = nblk / 2, rounds = nblk - 1;
    const int inner = (nblk == 2) ? 40 : 1;
    double* cur = w.buf0.p;
    double* nxt = w.buf1.p;
    {
        dim3 grid((unsigned)cdiv(m, 256), (unsigned)mp, (unsigned)nb);
        svd_init_batched_kernel<<<grid, 256, 0, st>>>(A, lda, m, mp, cur, ldw);
        LRN_CHECK_LAUNCH();
    }
    const int splits = w.splits, Kc = w.Kc, np = pairs * nb;
    int sweeps = 0;
    auto sweep_body = [&]() {
        LRN_CUDA(cudaMemsetAsync(w.offmax.p, 0, sizeof(double), st));
        for (int r = 0; r < rounds; r++) {
            GemmParams g;
            g.A = cur; g.B = cur; g.C = w.gram.p;
            g.transA = true; g.transB = false;
            g.M = 64; g.N = 64; g.K = Kc; g.lda = ldw; g.ldb = ldw; g.ldc = 64;
            g.batch = np; g.sA = (long long)64 * ldw; g.sB = (long long)64 * ldw; g.sC = (long long)splits * 4096;
            g.batch2 = splits; g.sA2 = Kc; g.sB2 = Kc; g.sC2 = 4096;
            g.K_last = m - (splits - 1) * Kc;
            gemm(g, st);
            EigSmallParams e;
            e.A = w.gram.p; e.lda = 64; e.sA = (long long)splits * 4096; e.nparts = splits; e.sPart = 4096; e.n = 64;
            e.V = w.rot.p; e.ldv = 64; e.sV = 4096; e.relative = 1; e.offmax = w.offmax.p; e.batch = np;
            e.max_sweeps = inner;
            e.cross_only = (nblk > 2 && r > 0) ? 1 : 0;
            jacobi_eig_small(e, st);
            GemmParams u;
            u.A = cur; u.B = w.rot.p; u.C = nxt;
            u.M = m; u.N = 64; u.K = 64; u.lda = ldw; u.ldb = 64; u.ldc = ldw;
            u.batch = np; u.sA = (long long)64 * ldw; u.sB = 4096;
            u.cblkmap = w.slotmap.p;
            gemm(u, st);
            std::swap(cur, nxt);
        }
    };
    auto sweep_advance = [&]() {
        if (rounds & 1) std::swap(cur, nxt);
    };
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
        run_sweep(w.graphs, (rounds & 1) ? (sweep & 1) : 0, st, sweep_body, sweep_advance);
        sweeps++;
        double off = 0.0;
        LRN_CUDA(cudaMemcpyAsync(&off, w.offmax.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        LRN_CUDA(cudaStreamSynchronize(st));
        if (off <= tol) break;
    }
    colnorm_kernel<<<(unsigned)cdiv((long long)mp * nb * 32, 256), 256, 0, st>>>(cur, ldw, m, mp * nb, w.sv.p);
    LRN_CHECK_LAUNCH();
    rank_desc_batched_kernel<<<dim3((unsigned)cdiv(mp, 256), (unsigned)nb), 256, 0, st>>>(w.sv.p, mp, w.perm.p);
    LRN_CHECK_LAUNCH();
    svd_gather_batched_kernel<<<dim3((unsigned)cdiv(m, 256), (unsigned)m, (unsigned)nb), 256, 0, st>>>(cur, ldw, m, mp, w.perm.p, w.sv.p,
                                                                                                     UD, ldu, sigma);
    LRN_CHECK_LAUNCH();
    return sweeps;
}

// ---------------------------------------------------------------------------------------------------------------
// host tridiagonal QL
// ---------------------------------------------------------------------------------------------------------------
bool tridiag_ql(int n, double* d, double* e_in, double* Z, double* zlast) {
    if (n <= 0) return true;
    std::vector<double> e(n, 0.0);
    for (int i = 0; i + 1 < n; i++) e[i] = e_in[i];
    std::vector<double> zl;
    if (Z) {
        for (int i = 0; i < n * n; i++) Z[i] = 0.0;
        for (int i = 0; i < n; i++) Z[i * n + i] = 1.0;
    } else if (zlast) {
        zl.assign(n, 0.0);
        zl[n - 1] = 1.0;
    }
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (iter++ == 80) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i], b = c * e[i];
                    r = std::hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    if (Z) {
                        for (int k = 0; k < n; k++) {
                            double fz = Z[k * n + i + 1];
                            Z[k * n + i + 1] = s * Z[k * n + i] + c * fz;
                            Z[k * n + i] = c * Z[k * n + i] - s * fz;
                        }
                    } else if (zlast) {
                        double fz = zl[i + 1];
                        zl[i + 1] = s * zl[i] + c * fz;
                        zl[i] = c * zl[i] - s * fz;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    // ascending sort (selection; n is small)
    for (int i = 0; i < n - 1; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++)
            if (d[j] < d[k]) k = j;
        if (k != i) {
            std::swap(d[i], d[k]);
            if (Z) for (int r = 0; r < n; r++) std::swap(Z[r * n + i], Z[r * n + k]);
            else if (zlast) std::swap(zl[i], zl[k]);
        }
    }
    if (zlast) {
        if (Z) for (int j = 0; j < n; j++) zlast[j] = Z[(n - 1) * n + j];
        else for (int j = 0; j < n; j++) zlast[j] = zl[j];
    }
    return true;
}

void LanczosWork::ensure(int m_, int kmax_) {
    if (m_ > m || kmax_ > kmax || !Q.p) {
        m = std::max(m, m_);
        kmax = std::max(kmax, kmax_);
        Q.alloc((size_t)pad_ld(m) * (kmax + 1));
        w.alloc(pad_ld(m));
        c.alloc((size_t)(kmax + 2) * RSPLIT);
        ab.alloc((size_t)2 * kmax + 2);
        S.alloc((size_t)(kmax + 1) * 64);
    }
    if (!scal.p) scal.alloc(8);
    if (!h_scal) LRN_CUDA(cudaMallocHost(&h_scal, 8 * sizeof(double)));
}
LanczosWork::~LanczosWork() {
    if (h_scal) cudaFreeHost(h_scal);
}

LanczosResult lanczos_extreme(const double* T, int m, int ld, int want, int nev_top, double* top_vals_host, double* top_vecs,
                              int ldv, double tol, LanczosWork& w, cudaStream_t st, int kmax_cap) {
    LanczosResult res;
    LRN_REQUIRE(nev_top >= 0 && nev_top <= 32 && nev_top < m, "nev_top out of range");
    if (m <= EN) {
        // direct: all eigenpairs by Jacobi in one CTA
        w.ensure(EN, EN);
        EigSmallParams e;
        e.A = T; e.lda = ld; e.n = m; e.evals = w.c.p; e.sort_desc = 1; e.batch = 1;
        if (nev_top > 0) { e.V = w.Q.p; e.ldv = pad_ld(w.m); }
        jacobi_eig_small(e, st);
        std::vector<double> ev(m);
        LRN_CUDA(cudaMemcpyAsync(ev.data(), w.c.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
        LRN_CUDA(cudaStreamSynchronize(st));
        res.lmax = ev[0]; res.lmin = ev[m - 1]; res.iters = 0; res.converged = true;
        for (int t = 0; t < nev_top; t++) {
            // ascending tail order: top_vals[0] is the smallest of the nev_top largest
            int src = nev_top - 1 - t;          // column index in descending order
            top_vals_host[t] = ev[src];
            LRN_CUDA(cudaMemcpyAsync(top_vecs + (size_t)t * ldv, w.Q.p + (size_t)src * pad_ld(w.m), m * sizeof(double),
                                     cudaMemcpyDeviceToDevice, st));
        }
        return res;
    }
    const int kmax = std::min(m, std::max(std::max(nev_top + 2, 4), kmax_cap));
    w.ensure(m, std::min(m, 500));
    const int ldq = pad_ld(w.m);
    double* Q = w.Q.p;
    std::vector<double> alpha, beta;     // beta[j] couples q_j and q_{j+1}
    lanczos_init_kernel<<<1, 1024, 0, st>>>(Q, m);
    LRN_CHECK_LAUNCH();
    // alpha_j, beta_j stay on the device (w.ab: alpha at [j], beta at [kmax + j]); the host only looks at them at the
    // check points, so the iterations in between are enqueued back to back without a stream synchronisation
    double beta_prev = 0.0;
    int next_check = 8;
    std::vector<double> d, e, zl;
    double scale = 0.0;
    int k = 0;
    bool done = false;
    double* dal = w.ab.p;
    double* dbe = w.ab.p + kmax;
    while (!done) {
        const int j = k;
        double* qj = Q + (size_t)j * ldq;
        gemv_t_kernel<<<(unsigned)cdiv((long long)m * 32, 256), 256, 0, st>>>(T, ld, m, m, qj, w.w.p);
        LRN_CHECK_LAUNCH();
        lanczos_alpha_kernel<<<1, 1024, 0, st>>>(w.w.p, qj, j > 0 ? Q + (size_t)(j - 1) * ldq : nullptr, j > 0 ? dbe + j - 1 : nullptr,
                                                 m, dal + j);
        LRN_CHECK_LAUNCH();
        for (int pass = 0; pass < 2; pass++) {   // full re-orthogonalisation, classical Gram-Schmidt twice
            gemv_t_split_kernel<<<dim3((unsigned)(j + 1), RSPLIT), 128, 0, st>>>(Q, ldq, m, w.w.p, w.c.p);
            LRN_CHECK_LAUNCH();
            gemv_n_sub_split_kernel<<<(unsigned)cdiv(m, 128), 128, (size_t)(j + 1) * sizeof(double), st>>>(Q, ldq, m, j + 1, w.c.p,
                                                                                                           w.w.p);
            LRN_CHECK_LAUNCH();
        }
        lanczos_beta_kernel<<<1, 1024, 0, st>>>(w.w.p, Q + (size_t)(j + 1) * ldq, m, dbe + j);
        LRN_CHECK_LAUNCH();
        k++;
        if (k < kmax && k < next_check) continue;
        alpha.resize(k);
        beta.resize(k);
        LRN_CUDA(cudaMemcpyAsync(alpha.data(), dal, k * sizeof(double), cudaMemcpyDeviceToHost, st));
        LRN_CUDA(cudaMemcpyAsync(beta.data(), dbe, k * sizeof(double), cudaMemcpyDeviceToHost, st));
        LRN_CUDA(cudaStreamSynchronize(st));
        // an exhausted Krylov space (beta_j ~ 0) ends the recurrence at the first such j (later vectors are zero / noise)
        scale = 0.0;
        bool breakdown = false;
        for (int t = 0; t < k; t++) {
            scale = std::max(scale, std::fabs(alpha[t]) + beta[t]);
            if (!(beta[t] > 1e-13 * scale)) { k = t + 1; alpha.resize(k); beta.resize(k); breakdown = true; break; }
        }
        beta_prev = beta[k - 1];
        {
            d = alpha;
            e.assign(beta.begin(), beta.end() - 1);
            e.push_back(0.0);
            zl.assign(k, 0.0);
            tridiag_ql(k, d.data(), e.data(), nullptr, zl.data());
            double sc = std::max(std::fabs(d[0]), std::fabs(d[k - 1]));
            if (sc == 0.0) sc = 1.0;
            bool ok = true;
            if (want & 1) ok = ok && (std::fabs(beta_prev * zl[0]) <= tol * sc);
            if (want & 2)
                for (int t = 0; t < std::max(nev_top, 1) && t < k; t++) ok = ok && (std::fabs(beta_prev * zl[k - 1 - t]) <= tol * sc);
            res.lmin = d[0];
            res.lmax = d[k - 1];
            res.resid_min = std::fabs(beta_prev * zl[0]);
            res.resid_max = std::fabs(beta_prev * zl[k - 1]);
            if (ok || breakdown || k >= kmax) {
                res.converged = ok || breakdown;
                done = true;
            }
            next_check = k + std::max(4, k / 4);
        }
    }
    res.iters = k;
    if (nev_top > 0) {
        std::vector<double> Z((size_t)k * k);
        d = alpha;
        e.assign(beta.begin(), beta.end() - 1);
        e.push_back(0.0);
        tridiag_ql(k, d.data(), e.data(), Z.data(), nullptr);
        const int nv = std::min(nev_top, k);
        std::vector<double> S((size_t)k * nv);
        for (int t = 0; t < nv; t++) {
            int src = k - nv + t;              // ascending tail
            top_vals_host[t] = d[src];
            for (int i = 0; i < k; i++) S[(size_t)t * k + i] = Z[(size_t)i * k + src];
        }
        if ((size_t)k * nv > w.S.n) w.S.alloc((size_t)k * nv);
        w.S.upload(S.data(), S.size(), st);
        gemm_nn(st, m, nv, k, 1.0, Q, ldq, w.S.p, k, 0.0, top_vecs, ldv);
        LRN_CUDA(cudaStreamSynchronize(st));
    }
    return res;
}

}  // namespace lrn
