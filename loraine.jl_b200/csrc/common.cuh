// Shared helpers for the loraine_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <atomic>

namespace lrn {

extern std::atomic<long long> g_kernel_launches;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define LRN_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            char buf__[512];                                                                        \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,             \
                     cudaGetErrorString(e__));                                                      \
            throw lrn::CudaError(buf__);                                                            \
        }                                                                                           \
    } while (0)

// every kernel launch site is followed by this check; it also feeds lrn_kernel_launches()
#define LRN_CHECK_LAUNCH()                                                   \
    do {                                                                     \
        LRN_CUDA(cudaGetLastError());                                        \
        lrn::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);      \
    } while (0)

#define LRN_REQUIRE(cond, msg)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            char buf__[512];                                                                        \
            snprintf(buf__, sizeof buf__, "%s:%d: requirement failed: %s (%s)", __FILE__, __LINE__, \
                     #cond, msg);                                                                   \
            throw std::invalid_argument(buf__);                                                     \
        }                                                                                           \
    } while (0)

// Runs `f` once per CUDA device (function attributes such as the dynamic shared-memory opt-in are per device; a process may
// drive several devices, one handle each).  Thread-safe for concurrent callers on different devices.
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    template <typename F>
    void run(F&& f) {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        if (mask.load(std::memory_order_acquire) & bit) return;
        f();
        mask.fetch_or(bit, std::memory_order_release);
    }
};

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// Leading dimension used for every dense device matrix: multiple of 8 doubles (64 B) so that every column
// start is 16-byte aligned (cp.async 16 B chunks) and sector aligned.
static inline int pad_ld(int m) { return round_up(m < 1 ? 1 : m, 8); }

// RAII device buffer (zero-initialised so that padding rows never hold NaNs).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return;
        LRN_CUDA(cudaMalloc(&p, count * sizeof(T)));
        LRN_CUDA(cudaMemset(p, 0, count * sizeof(T)));
        // the memset runs on the legacy stream, the library works on non-blocking streams: order them explicitly
        LRN_CUDA(cudaDeviceSynchronize());
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    void upload(const T* h, size_t count, cudaStream_t s = 0) {
        if (count > n) alloc(count);
        if (count) LRN_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& h, cudaStream_t s = 0) { upload(h.data(), h.size(), s); }
};

// Dense column-major device matrix with padded leading dimension.
struct DMat {
    DevBuf<double> buf;
    int rows = 0, cols = 0, ld = 0;
    DMat() = default;
    DMat(int r, int c) { init(r, c); }
    void init(int r, int c) {
        rows = r; cols = c; ld = pad_ld(r);
        buf.alloc((size_t)ld * (size_t)(c < 1 ? 1 : c));
    }
    double* p() const { return buf.p; }
    size_t bytes() const { return (size_t)ld * cols * sizeof(double); }
};

// ---- device-side reduction helpers -------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum; result valid in thread 0 (and broadcast to all when `bcast`). blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* sh /* >= 32 doubles */) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    double r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0) r = warp_sum(r);
    if (threadIdx.x == 0) sh[0] = r;
    __syncthreads();
    return sh[0];
}
#endif

}  // namespace lrn
