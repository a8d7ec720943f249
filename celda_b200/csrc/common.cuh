// Shared host-side plumbing for libcelda_cuda: error reporting, device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace celda {

inline thread_local char g_err[1024] = "";
inline std::atomic<long long> g_launches{0};

inline int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return 1;
}

#define CELDA_CUDA_OK(expr)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return celda::fail("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #expr); \
    } while (0)

#define CELDA_TRY(expr)          \
    do {                         \
        int rc__ = (expr);       \
        if (rc__) return rc__;   \
    } while (0)

inline void count_launch(long long n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// RAII device buffer (freed on scope exit; stateless entry points rely on this to release before returning errors).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) count = 1;
        CELDA_CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
        return 0;
    }
    int upload(const T* host, size_t count, cudaStream_t st = 0) {
        CELDA_TRY(alloc(count));
        if (count) CELDA_CUDA_OK(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
        return 0;
    }
    int download(T* host, size_t count, cudaStream_t st = 0) const {
        if (count) CELDA_CUDA_OK(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, st));
        return 0;
    }
    int zero(cudaStream_t st = 0) {
        if (n) CELDA_CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(T), st));
        return 0;
    }
};

inline int num_sms(int device) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    return v > 0 ? v : 148;
}

}  // namespace celda
