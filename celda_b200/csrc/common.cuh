// Shared host-side plumbing for libcelda_cuda: error reporting, device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
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
    cudaDeviceSynchronize();  // the caller's buffers return to the cache (DevCache) next: nothing may still be using them
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

// Device memory cache.  The split heuristics and the 'split' initialisation create and destroy hundreds of short-lived
// chain handles on sub-matrices (R/split_clusters.R, R/initialize_clusters.R); cudaMalloc / cudaFree (which also
// synchronises the device) would dominate them.  Released blocks are kept per device and handed out again to requests
// of a similar size (at most 25 % larger than needed); everything is returned to the driver when an allocation fails
// or the cache exceeds its cap.
struct DevCache {
    std::mutex mu;
    std::multimap<size_t, void*> blocks[16];
    size_t held[16] = {0};
    static constexpr size_t kCap = (size_t)48 << 30;
    static size_t round_up(size_t b) { return b < 4096 ? 4096 : (b + 4095) & ~(size_t)4095; }
    void* take(int dev, size_t bytes, size_t* got) {
        std::lock_guard<std::mutex> g(mu);
        auto& m = blocks[dev & 15];
        auto it = m.lower_bound(bytes);
        if (it == m.end() || it->first > bytes + bytes / 4) return nullptr;
        void* p = it->second;
        *got = it->first;
        held[dev & 15] -= it->first;
        m.erase(it);
        return p;
    }
    bool give(int dev, void* p, size_t bytes) {
        std::lock_guard<std::mutex> g(mu);
        if (held[dev & 15] + bytes > kCap) return false;
        blocks[dev & 15].emplace(bytes, p);
        held[dev & 15] += bytes;
        return true;
    }
    void flush(int dev) {
        std::lock_guard<std::mutex> g(mu);
        for (auto& kv : blocks[dev & 15]) cudaFree(kv.second);
        blocks[dev & 15].clear();
        held[dev & 15] = 0;
    }
};
inline DevCache g_devcache;

// RAII device buffer (released on scope exit; stateless entry points rely on this to release before returning errors).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    size_t cap_bytes = 0;
    int dev = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p && !g_devcache.give(dev, p, cap_bytes)) cudaFree(p);
        p = nullptr;
        n = 0;
        cap_bytes = 0;
    }
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) count = 1;
        cudaGetDevice(&dev);
        const size_t want = DevCache::round_up(count * sizeof(T));
        if (void* c = g_devcache.take(dev, want, &cap_bytes)) {
            p = static_cast<T*>(c);
            return 0;
        }
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {  // give the cached memory back to the driver and try once more
            cudaGetLastError();
            g_devcache.flush(dev);
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            p = nullptr;
            return celda::fail("CUDA error %s allocating %zu bytes", cudaGetErrorString(e), want);
        }
        cap_bytes = want;
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
