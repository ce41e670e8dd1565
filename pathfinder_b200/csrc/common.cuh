// pathfinder_b200/csrc/common.cuh — shared device/host plumbing for the CUDA backend.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

struct PFScene;

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

// Set by PFSceneBuildAndRenderCuda around PFSceneBuild: the scene's segment arrays outlive the frame
// (they are only rewritten after the scene has waited for the renderer's copies), so UploadSceneD3D11
// may carry payload_persists = 1.
extern thread_local bool g_scene_payload_persists;
// Tile rows [y0, y1) of the renderer the scene is being built for (y1 <= y0: the whole frame); set by
// PFSceneBuildAndRenderCuda around PFSceneBuild.
extern thread_local int32_t g_scene_strip[2];
// Records, on the renderer's stream, the point after which the scene may rewrite the segment arrays it
// lent to the renderer (scene.cpp waits for it before the next rebuild).
void scene_note_borrowed(::PFScene *scene, cudaStream_t stream, int device);

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Programmatic dependent launch (sm_90+): a kernel launched with launch_chained is set up while the kernel before it
// on the stream is still running and its blocks are scheduled as that kernel's blocks exit — they sit in chain_wait()
// until it has completed and its writes are visible — so the launch latency of one stage hides behind the stage before
// it. Every kernel of the frame's chain starts with chain_wait() (a no-op when it was launched the ordinary way). No
// kernel triggers its dependents early (griddepcontrol.launch_dependents at the top of a kernel was measured: the
// waiting blocks of the next stage then hold SM resources the other scenes' streams could use — isolated frames the
// same, the three-scene step 4 % slower). PF_CUDA_NO_PDL=1 in the environment launches the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool chained_launches_enabled() {
    static const bool on = [] {
        const char *e = getenv("PF_CUDA_NO_PDL");
        return !(e && e[0] == '1');
    }();
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_chained(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = 0, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = chained_launches_enabled() ? 1 : 0;
    PF_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}
#endif

// PF_HOST_TIMING=1: per-phase host times of the scene / batch build, printed to stderr.
struct LapTimer {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        static const bool on = getenv("PF_HOST_TIMING") != nullptr;
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "  %s: %.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// What a spinning host thread does between two looks at a flag.
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
}

// Persistent host worker threads (the reference parallelises its CPU-side scene work with Rayon's
// global pool, renderer/src/concurrent/rayon.rs:17-24). Spawning threads per call costs more than
// the per-frame work items themselves (a 100k-path batch build is ~1 ms), so the workers are kept
// and woken through a condition variable; they spin for a few tens of microseconds first because the
// build runs several parallel phases back to back.
class WorkerPool {
  public:
    static WorkerPool &instance() {
        static WorkerPool pool;
        return pool;
    }
    size_t width() const { return workers_.size() + 1; }

    // Runs job(chunk) for chunk in [0, n_chunks) on the workers and the calling thread.
    template <typename F>
    void run(size_t n_chunks, F &&job) {
        if (n_chunks == 0) return;
        if (n_chunks > MAX_CHUNKS) throw Error(1, "WorkerPool: too many chunks");
        std::lock_guard<std::mutex> serial(run_mutex_); // one parallel region at a time
        struct Thunk {
            static void call(void *ctx, size_t i) { (*static_cast<typename std::remove_reference<F>::type *>(ctx))(i); }
        };
        fn_ = &Thunk::call;
        ctx_ = &job;
        remaining_.store(n_chunks, std::memory_order_relaxed);
        const uint64_t gen = ((state_.load(std::memory_order_relaxed) >> GEN_SHIFT) + 1) & GEN_MASK;
        {
            std::lock_guard<std::mutex> lock(mutex_); // sleeping workers re-check the state under this mutex
            state_.store(gen << GEN_SHIFT | (uint64_t)n_chunks << COUNT_SHIFT, std::memory_order_release);
        }
        wake_.notify_all();
        work(gen);
        // Wait for chunks still running on workers.
        for (int spin = 0; remaining_.load(std::memory_order_acquire) != 0; spin++) {
            if (spin > 2000) std::this_thread::yield();
            else cpu_relax();
        }
    }

  private:
    // state_ = generation | chunk count | next chunk. A worker claims a chunk with a compare-exchange
    // on the whole word, so a worker that is late for job g can neither run nor skip a chunk of job g+1.
    static constexpr int COUNT_SHIFT = 20, GEN_SHIFT = 40;
    static constexpr uint64_t INDEX_MASK = (1ull << COUNT_SHIFT) - 1, GEN_MASK = (1ull << 24) - 1;
    static constexpr size_t MAX_CHUNKS = INDEX_MASK;

    WorkerPool() {
        // PF_HOST_THREADS caps the pool (several renderer processes on one host should share its cores).
        unsigned hw = std::thread::hardware_concurrency();
        size_t n = std::min<size_t>(hw ? hw : 1, 16);
        if (const char *env = getenv("PF_HOST_THREADS")) n = std::max<size_t>(1, std::min<size_t>(n, (size_t)atoi(env)));
        for (size_t i = 1; i < n; i++) workers_.emplace_back([this]() { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            stop_.store(true, std::memory_order_release);
        }
        wake_.notify_all();
        for (auto &w : workers_) w.join();
    }
    void work(uint64_t gen) {
        uint64_t v = state_.load(std::memory_order_acquire);
        for (;;) {
            const uint64_t index = v & INDEX_MASK, count = (v >> COUNT_SHIFT) & INDEX_MASK;
            if ((v >> GEN_SHIFT) != gen || index >= count) return;
            if (!state_.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel, std::memory_order_acquire)) continue;
            fn_(ctx_, (size_t)index); // job `gen` cannot end before this chunk reports, so fn_/ctx_ are its own
            remaining_.fetch_sub(1, std::memory_order_acq_rel);
            v = state_.load(std::memory_order_acquire);
        }
    }
    void loop() {
        uint64_t seen = 0;
        auto changed = [&]() {
            return (state_.load(std::memory_order_acquire) >> GEN_SHIFT) != seen || stop_.load(std::memory_order_acquire);
        };
        for (;;) {
            // Spin for a short while (phases follow each other within microseconds), then sleep. The spin is a
            // thousand PAUSEs (tens of microseconds), not a bare load loop ten times as long: a busy spinner takes
            // issue slots from its hyper-thread sibling, and when the renderer processes of a multi-GPU job
            // oversubscribe the host's cores it burns time slices that the thread it waits for needs. (Eight builders
            // of two threads on eight noisy cores: the slowest one's step 5.2 -> 4.9 ms, the median 4.1 -> 3.9 ms over
            // six interleaved rounds — inside the noise of that machine, never slower.)
            bool woke = false;
            for (int spin = 0; spin < 1000 && !woke; spin++) {
                woke = changed();
                if (!woke) cpu_relax();
            }
            if (!woke) {
                std::unique_lock<std::mutex> lock(mutex_);
                wake_.wait(lock, changed);
            }
            if (stop_.load(std::memory_order_acquire)) return;
            seen = state_.load(std::memory_order_acquire) >> GEN_SHIFT;
            work(seen);
        }
    }

    std::vector<std::thread> workers_;
    std::mutex mutex_, run_mutex_;
    std::condition_variable wake_;
    void (*fn_)(void *, size_t) = nullptr;
    void *ctx_ = nullptr;
    std::atomic<uint64_t> state_{0};
    std::atomic<size_t> remaining_{0};
    std::atomic<bool> stop_{false};
};

// Fixed chunking of [0, n): at most one chunk per pool thread, at least `grain` items each. Passes
// that share (n, grain) see the same chunk boundaries, so a per-chunk partial sum of one pass can seed
// the running offsets of the next (two-level scan).
inline size_t chunk_count(size_t n, size_t grain) {
    if (grain == 0) grain = 1;
    return std::max<size_t>(1, std::min(WorkerPool::instance().width(), n / grain));
}

// Runs f(chunk, begin, end) for every chunk of [0, n) on the worker pool; a single chunk runs inline.
template <typename F>
inline void parallel_chunks(size_t n, size_t chunks, F &&f) {
    if (chunks <= 1) {
        f((size_t)0, (size_t)0, n);
        return;
    }
    const size_t per = (n + chunks - 1) / chunks;
    WorkerPool::instance().run(chunks, [&](size_t c) {
        const size_t b0 = std::min(n, c * per), e0 = std::min(n, b0 + per);
        f(c, b0, e0); // possibly empty: every chunk reports, so per-chunk results are always written
    });
}

// Runs f(begin, end) over [0, n) split into ranges of at least `grain` items. Small inputs run inline.
template <typename F>
inline void parallel_ranges(size_t n, size_t grain, F &&f) {
    parallel_chunks(n, chunk_count(n, grain), [&](size_t, size_t b0, size_t e0) {
        if (b0 < e0 || n == 0) f(b0, e0);
    });
}

} // namespace pf
