// pathfinder_b200/csrc/scan.cuh — device-wide exclusive prefix sum (u32), hand-written.
//
// Count -> scan -> emit replaces the reference's atomicAdd allocation + host read-back + retry
// (shaders/d3d11/dice.cs.glsl:204-216, bin.cs.glsl:100-117; renderer/src/gpu/d3d11/renderer.rs
// :218-229,338-353,463-499): every stage first counts its outputs, the scan turns counts into
// deterministic offsets, and the stage then writes in place.
//
// One launch, one pass over the input ("decoupled look-back"): every block takes a ticket, reduces
// its 4096-element tile, publishes the tile aggregate, looks back over its predecessors' published
// aggregates / inclusive prefixes until it knows its own exclusive prefix, publishes its inclusive
// prefix and writes the scanned tile. Status words carry an epoch so they never need clearing; the
// epoch and the ticket counter live in device memory and are advanced by the last tile, which keeps
// the launch replayable from a CUDA graph. The input is a functor so producers can be fused
// (e.g. "low 24 bits of the tile word").
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

// Device-resident state shared by all scans on one stream (they run back to back, never
// concurrently): control[0] = ticket counter, control[1] = epoch; one status word per tile.
struct ScanState {
    uint32_t *control;
    unsigned long long *status; // epoch << 34 | flag << 32 | value
};

constexpr unsigned long long SCAN_FLAG_AGGREGATE = 1ull, SCAN_FLAG_PREFIX = 2ull;

__device__ __forceinline__ unsigned long long scan_pack(uint32_t epoch, unsigned long long flag, uint32_t value) {
    return ((unsigned long long)(epoch & 0x3fffffffu) << 34) | (flag << 32) | (unsigned long long)value;
}

// `n_dev` (optional): the element count lives in device memory (a previous stage's total); `n` is
// then the host-side bound the grid was sized with, and the kernel uses min(*n_dev, n).
template <typename InFn>
__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan(InFn in, uint32_t n, const uint32_t *__restrict__ n_dev, uint32_t *__restrict__ out,
           uint32_t *__restrict__ total_out, ScanState st) {
    __shared__ uint32_t smem[SCAN_THREADS / 32 + 1];
    __shared__ uint32_t s_tile, s_epoch, s_prefix;
    chain_wait(); // (programmatic dependent launch, see common.cuh)
    if (n_dev) n = min(n, *n_dev);
    if (threadIdx.x == 0) {
        s_epoch = *reinterpret_cast<volatile uint32_t *>(st.control + 1) + 1; // read before the ticket
        s_tile = atomicAdd(st.control, 1u);
    }
    __syncthreads();
    const uint32_t tile = s_tile, epoch = s_epoch;
    const uint32_t n_tiles = gridDim.x;

    // Load and reduce the tile (items stay in registers).
    const size_t base = (size_t)tile * SCAN_TILE;
    uint32_t v[4][4];
    uint32_t round_sum[4];
    uint32_t sum = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        size_t i0 = base + (size_t)r * (SCAN_THREADS * 4) + (size_t)threadIdx.x * 4;
        round_sum[r] = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[r][k] = (i0 + k < n) ? in((uint32_t)(i0 + k)) : 0;
            round_sum[r] += v[r][k];
        }
        sum += round_sum[r];
    }
    uint32_t aggregate;
    block_exclusive_scan(sum, smem, &aggregate);

    // Publish, look back, publish again.
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        volatile unsigned long long *status = st.status;
        if (tile == 0) {
            if (lane == 0) {
                status[0] = scan_pack(epoch, SCAN_FLAG_PREFIX, aggregate);
                s_prefix = 0;
            }
        } else {
            if (lane == 0) status[tile] = scan_pack(epoch, SCAN_FLAG_AGGREGATE, aggregate);
            uint32_t exclusive = 0;
            int look = (int)tile - 1;
            for (;;) {
                int idx = look - (int)lane;
                unsigned long long w;
                unsigned long long flag;
                do {
                    w = idx >= 0 ? status[idx] : scan_pack(epoch, SCAN_FLAG_PREFIX, 0);
                    flag = ((w >> 34) == (unsigned long long)(epoch & 0x3fffffffu)) ? ((w >> 32) & 3ull) : 0ull;
                } while (__any_sync(0xffffffffu, flag == 0ull));
                unsigned prefix_lanes = __ballot_sync(0xffffffffu, flag == SCAN_FLAG_PREFIX);
                uint32_t val = (uint32_t)w;
                if (prefix_lanes) {
                    unsigned first = __ffs(prefix_lanes) - 1; // nearest predecessor with an inclusive prefix
                    val = lane <= first ? val : 0;
                }
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                exclusive += val;
                if (prefix_lanes) break;
                look -= 32;
            }
            if (lane == 0) {
                status[tile] = scan_pack(epoch, SCAN_FLAG_PREFIX, exclusive + aggregate);
                s_prefix = exclusive;
            }
        }
    }
    __syncthreads();
    uint32_t carry = s_prefix;
    if (tile == n_tiles - 1 && threadIdx.x == 0) {
        if (total_out) *total_out = carry + aggregate;
        // Every other tile already holds its ticket and epoch: advance both for the next scan.
        st.control[1] = epoch;
        st.control[0] = 0;
    }

    // Scan the tile and write it out.
#pragma unroll
    for (int r = 0; r < 4; r++) {
        size_t i0 = base + (size_t)r * (SCAN_THREADS * 4) + (size_t)threadIdx.x * 4;
        uint32_t total;
        uint32_t excl = block_exclusive_scan(round_sum[r], smem, &total) + carry;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) out[i0 + k] = excl;
            excl += v[r][k];
        }
        carry += total;
    }
}

struct ScanScratch {
    DeviceBuffer<uint32_t> control;
    DeviceBuffer<unsigned long long> status;
    bool initialised = false;
};

// out[i] = sum_{j<i} in(j) for i in [0, n); *total_out (device) = sum of all. `out` may alias
// the array `in` reads if in(i) reads element i alone (a tile is read completely before it is
// written, and tiles are disjoint). Returns the number of kernels launched.
template <typename InFn>
inline int exclusive_scan(InFn in, uint32_t *out, uint32_t n, uint32_t *total_out, ScanScratch &scratch,
                          cudaStream_t stream, const uint32_t *n_dev = nullptr) {
    if (n == 0) {
        if (total_out) PF_CUDA_CHECK(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream));
        return 0;
    }
    unsigned n_blocks = div_up(n, SCAN_TILE);
    if (!scratch.initialised) {
        scratch.control.ensure(2);
        PF_CUDA_CHECK(cudaMemsetAsync(scratch.control.ptr, 0, 2 * sizeof(uint32_t), stream));
        scratch.initialised = true;
    }
    if (n_blocks > scratch.status.capacity) {
        // The old buffer may still be in use by an earlier scan on this stream.
        PF_CUDA_CHECK(cudaStreamSynchronize(stream));
        scratch.status.ensure(n_blocks, 1.5);
        PF_CUDA_CHECK(cudaMemsetAsync(scratch.status.ptr, 0, scratch.status.capacity * sizeof(unsigned long long), stream));
    }
    ScanState st{scratch.control.ptr, scratch.status.ptr};
    launch_chained(k_scan<InFn>, n_blocks, SCAN_THREADS, stream, in, n, n_dev, out, total_out, st);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
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
