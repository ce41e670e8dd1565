// pathfinder_b200/csrc/radix_sort.cuh — stable LSD radix sort of (u32 key, u32 value) pairs,
// hand-written: 8 bits per pass, per-block digit histograms -> device scan -> stable scatter.
//
// Used by the "sort" stage: tile-list entries are generated in tile-index order (= draw order,
// tiles are allocated path by path) and sorted by framebuffer tile, so each framebuffer tile ends
// with its tiles in painter's order. Replaces the per-framebuffer-tile linked-list insertion sort
// of shaders/d3d11/sort.cs.glsl:60-95.
//
// Stability inside a block comes from warp-ordered ranking: warp w owns a contiguous slice of the
// block's tile and walks it in rounds of 32 consecutive items; within a round lanes with the same
// digit are ranked with __match_any_sync.
#pragma once

#include "common.cuh"
#include "scan.cuh"

namespace pf {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16; // items per lane
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;
constexpr int RS_BINS = 256;

// hist[bin * n_blocks + block] = number of keys in the block's tile with that digit.
__global__ void __launch_bounds__(RS_THREADS)
    k_rs_histogram(const uint32_t *keys, uint32_t n, const uint32_t *__restrict__ n_dev, int shift,
                   uint32_t n_blocks, uint32_t *hist) {
    __shared__ uint32_t s_hist[RS_BINS];
    if (n_dev) n = min(n, *n_dev);
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; r++) {
        size_t i = base + (size_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[(keys[i] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
    k_rs_scatter(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                 uint32_t n, const uint32_t *__restrict__ n_dev, int shift, uint32_t n_blocks,
                 const uint32_t *hist_scanned) {
    // Per-warp digit counts, then turned into per-warp exclusive bases.
    __shared__ uint32_t s_warp_hist[RS_WARPS][RS_BINS];
    if (n_dev) n = min(n, *n_dev);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();

    // Warp w owns items [base + w*512, base + (w+1)*512): round r covers 32 consecutive items.
    const size_t warp_base = (size_t)blockIdx.x * RS_TILE + (size_t)warp * (32 * RS_ROUNDS);
    uint32_t key[RS_ROUNDS];
    uint32_t rank[RS_ROUNDS]; // rank among the warp's earlier items with the same digit
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        size_t i = warp_base + (size_t)r * 32 + lane;
        bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0xffffffffu;
        unsigned digit = (key[r] >> shift) & 0xff;
        // Invalid lanes take part in the match with an impossible 9-bit digit so they never rank.
        unsigned m = __match_any_sync(0xffffffffu, valid ? digit : 0x100u);
        unsigned earlier = __popc(m & ((1u << lane) - 1u));
        uint32_t before = valid ? s_warp_hist[warp][digit] : 0;
        rank[r] = before + earlier;
        __syncwarp();
        if (valid && earlier == 0) s_warp_hist[warp][digit] = before + __popc(m); // group leader
        __syncwarp();
    }
    __syncthreads();
    // Exclusive prefix over warps per digit + the block's global base for the digit.
    {
        unsigned digit = threadIdx.x; // RS_THREADS == RS_BINS
        uint32_t run = hist_scanned[(size_t)digit * n_blocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = s_warp_hist[w][digit];
            s_warp_hist[w][digit] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        size_t i = warp_base + (size_t)r * 32 + lane;
        if (i < n) {
            unsigned digit = (key[r] >> shift) & 0xff;
            uint32_t dst = s_warp_hist[warp][digit] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in[i];
        }
    }
}

struct RadixSortScratch {
    DeviceBuffer<uint32_t> hist;
    DeviceBuffer<uint32_t> keys_tmp, vals_tmp;
    ScanScratch scan;
};

// Sorts n pairs by the low `key_bits` bits of the key, stably. The result is left in
// (keys, vals) — an odd number of passes copies back through the temporaries. Returns launches.
// With `n_dev` the pair count is read on the device (min(*n_dev, n)); n is the host bound.
inline int radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t n, int key_bits, RadixSortScratch &s,
                            cudaStream_t stream, const uint32_t *n_dev = nullptr) {
    if (n == 0) return 0;
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    unsigned n_blocks = div_up(n, RS_TILE);
    s.hist.ensure((size_t)RS_BINS * n_blocks, 1.5);
    s.keys_tmp.ensure(n, 1.25);
    s.vals_tmp.ensure(n, 1.25);
    uint32_t *k_in = keys, *v_in = vals, *k_out = s.keys_tmp.ptr, *v_out = s.vals_tmp.ptr;
    int launches = 0;
    for (int p = 0; p < passes; p++) {
        int shift = 8 * p;
        k_rs_histogram<<<n_blocks, RS_THREADS, 0, stream>>>(k_in, n, n_dev, shift, n_blocks, s.hist.ptr);
        launches += 1;
        launches += exclusive_scan(LoadU32{s.hist.ptr}, s.hist.ptr, RS_BINS * n_blocks, nullptr, s.scan, stream);
        k_rs_scatter<<<n_blocks, RS_THREADS, 0, stream>>>(k_in, v_in, k_out, v_out, n, n_dev, shift, n_blocks, s.hist.ptr);
        launches += 1;
        uint32_t *t;
        t = k_in, k_in = k_out, k_out = t;
        t = v_in, v_in = v_out, v_out = t;
    }
    PF_CUDA_CHECK(cudaGetLastError());
    if (k_in != keys) {
        PF_CUDA_CHECK(cudaMemcpyAsync(keys, k_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
        PF_CUDA_CHECK(cudaMemcpyAsync(vals, v_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
    }
    return launches;
}

} // namespace pf
