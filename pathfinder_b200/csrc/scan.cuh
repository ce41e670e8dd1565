// pathfinder_b200/csrc/scan.cuh — device-wide exclusive prefix sum (u32), hand-written.
//
// Count -> scan -> emit replaces the reference's atomicAdd allocation + host read-back + retry
// (shaders/d3d11/dice.cs.glsl:204-216, bin.cs.glsl:100-117; renderer/src/gpu/d3d11/renderer.rs
// :218-229,338-353,463-499): every stage first counts its outputs, the scan turns counts into
// deterministic offsets, and the stage then writes in place.
//
// Three launches: per-block reduce, one-block scan of the block sums, per-block scan + offset.
// The input is a functor so that producers can be fused (e.g. "low 24 bits of the tile word").
#pragma once

#include "common.cuh"

namespace pf {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16; // per thread, as 4 rounds of 4 consecutive items
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += n;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread (SCAN_THREADS threads). Returns the
// exclusive prefix; *total receives the block sum. smem: 8 words + 1.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem, uint32_t *total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (SCAN_THREADS / 32) ? smem[lane] : 0;
        uint32_t winc = warp_inclusive_scan(w);
        if (lane < (SCAN_THREADS / 32)) smem[lane] = winc - w;
        if (lane == (SCAN_THREADS / 32) - 1) smem[SCAN_THREADS / 32] = winc;
    }
    __syncthreads();
    uint32_t res = inc - v + smem[warp];
    *total = smem[SCAN_THREADS / 32];
    __syncthreads();
    return res;
}

// `n_dev` (optional): the element count lives in device memory (a previous stage's total); `n` is
// then the host-side bound the grid was sized with, and the kernels use min(*n_dev, n).
template <typename InFn>
__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan_reduce(InFn in, uint32_t n, const uint32_t *__restrict__ n_dev, uint32_t *block_sums) {
    __shared__ uint32_t smem[SCAN_THREADS / 32 + 1];
    if (n_dev) n = min(n, *n_dev);
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t sum = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        size_t i0 = base + (size_t)r * (SCAN_THREADS * 4) + (size_t)threadIdx.x * 4;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i0 + k < n) sum += in((uint32_t)(i0 + k));
    }
    uint32_t total;
    block_exclusive_scan(sum, smem, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// One block: exclusive scan of the block sums in place, total to *total_out (+ optional copy).
__global__ void __launch_bounds__(1024) k_scan_block_sums(uint32_t *block_sums, uint32_t n_blocks,
                                                          uint32_t *total_out) {
    __shared__ uint32_t warp_sums[33];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n_blocks ? block_sums[i] : 0;
        uint32_t inc = warp_inclusive_scan(v);
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
            uint32_t winc = warp_inclusive_scan(w);
            warp_sums[lane] = winc - w;
            if (lane == 31) warp_sums[32] = winc;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        if (i < n_blocks) block_sums[i] = carry + warp_sums[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sums[32];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

template <typename InFn>
__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan_final(InFn in, uint32_t n, const uint32_t *__restrict__ n_dev, const uint32_t *block_offsets,
                 uint32_t *out) {
    __shared__ uint32_t smem[SCAN_THREADS / 32 + 1];
    if (n_dev) n = min(n, *n_dev);
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t carry = block_offsets[blockIdx.x];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        size_t i0 = base + (size_t)r * (SCAN_THREADS * 4) + (size_t)threadIdx.x * 4;
        uint32_t v[4];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k] = (i0 + k < n) ? in((uint32_t)(i0 + k)) : 0;
            sum += v[k];
        }
        uint32_t total;
        uint32_t excl = block_exclusive_scan(sum, smem, &total) + carry;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) out[i0 + k] = excl;
            excl += v[k];
        }
        carry += total;
    }
}

struct ScanScratch {
    DeviceBuffer<uint32_t> block_sums;
};

// out[i] = sum_{j<i} in(j) for i in [0, n); *total_out (device) = sum of all. `out` may alias
// the array `in` reads only if in(i) reads element i alone (each element is read before it is
// written by the same thread). Returns the number of kernels launched.
template <typename InFn>
inline int exclusive_scan(InFn in, uint32_t *out, uint32_t n, uint32_t *total_out, ScanScratch &scratch,
                          cudaStream_t stream, const uint32_t *n_dev = nullptr) {
    if (n == 0) {
        if (total_out) PF_CUDA_CHECK(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream));
        return 0;
    }
    unsigned n_blocks = div_up(n, SCAN_TILE);
    scratch.block_sums.ensure(n_blocks, 1.5);
    k_scan_reduce<<<n_blocks, SCAN_THREADS, 0, stream>>>(in, n, n_dev, scratch.block_sums.ptr);
    k_scan_block_sums<<<1, 1024, 0, stream>>>(scratch.block_sums.ptr, n_blocks, total_out);
    k_scan_final<<<n_blocks, SCAN_THREADS, 0, stream>>>(in, n, n_dev, scratch.block_sums.ptr, out);
    PF_CUDA_CHECK(cudaGetLastError());
    return 3;
}

struct LoadU32 {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i]; }
};
struct LoadLow24 {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i] & 0x00ffffffu; }
};
// Fill count of tiles that survived the z-cull (tile_fb != invalid), else 0.
struct LoadLiveCount {
    const uint32_t *word, *tile_fb;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        return tile_fb[i] != 0xffffffffu ? (word[i] & 0x00ffffffu) : 0u;
    }
};
struct LoadNotInvalid {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i] != 0xffffffffu ? 1u : 0u; }
};
struct LoadU8 {
    const uint8_t *p;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i]; }
};

} // namespace pf
