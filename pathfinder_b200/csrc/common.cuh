// pathfinder_b200/csrc/common.cuh — shared device/host plumbing for the CUDA backend.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace pf {

// Thrown by host code; the C ABI converts it into a status + PFCudaGetLastError() message (the
// reference panics instead, e.g. renderer/src/gpu/d3d11/renderer.rs:498-499).
struct Error : std::runtime_error {
    int status;
    Error(int status_, const std::string &msg) : std::runtime_error(msg), status(status_) {}
};

#define PF_CUDA_CHECK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            throw ::pf::Error(2, std::string(#expr) + " failed: " + cudaGetErrorString(err__) +   \
                                     " (" __FILE__ ":" + std::to_string(__LINE__) + ")");         \
        }                                                                                         \
    } while (0)

// Growable device buffer. Sizes are in elements of T. Growth reallocates (contents are not
// preserved unless asked) — the arena replaces the reference's GPUMemoryAllocator
// (gpu/src/allocator.rs) for this path.
template <typename T>
struct DeviceBuffer {
    T *ptr = nullptr;
    size_t capacity = 0;
    size_t *bytes_allocated = nullptr; // optional accounting

    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer() { release(); }

    void release() {
        if (ptr) {
            cudaFree(ptr);
            if (bytes_allocated) *bytes_allocated -= capacity * sizeof(T);
        }
        ptr = nullptr;
        capacity = 0;
    }
    // Ensures room for n elements; over-allocates by `slack` (e.g. 1.25) when growing.
    void ensure(size_t n, double slack = 1.0) {
        if (n <= capacity) return;
        size_t want = (size_t)((double)n * slack) + 16;
        release();
        PF_CUDA_CHECK(cudaMalloc((void **)&ptr, want * sizeof(T)));
        capacity = want;
        if (bytes_allocated) *bytes_allocated += capacity * sizeof(T);
    }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Runs f(begin, end) over [0, n) split across host threads (the reference parallelises its CPU-side
// scene work with Rayon, renderer/src/concurrent/rayon.rs:17-24). Small inputs run inline.
template <typename F>
inline void parallel_ranges(size_t n, size_t grain, F &&f) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t threads = std::min<size_t>(hw ? hw : 1, 16);
    threads = std::min(threads, n / (grain ? grain : 1));
    if (threads <= 1) {
        f((size_t)0, n);
        return;
    }
    std::vector<std::thread> workers;
    workers.reserve(threads);
    const size_t per = (n + threads - 1) / threads;
    for (size_t t = 0; t < threads; t++) {
        const size_t b0 = t * per, e0 = std::min(n, b0 + per);
        if (b0 >= e0) break;
        workers.emplace_back([&f, b0, e0]() { f(b0, e0); });
    }
    for (auto &w : workers) w.join();
}

} // namespace pf
