// pathfinder_b200/csrc/kernels.cu — the pipeline stages as hand-written CUDA for sm_100a.
//
// Compiled with -fmad=false and without --use_fast_math: the dice and bin stages must reproduce
// the reference CPU tiler's IEEE binary32 arithmetic operator by operator (one rounding each, no
// FMA contraction; '/' is IEEE division), because tile / fill / backdrop lists are compared
// bit-exactly. Where a fused multiply-add is wanted (coverage and
// compositing, compared at 1/255) it is written explicitly with fmaf().
//
// Reference files restated here (paths relative to the reference checkout):
//   dice       renderer/src/tiler.rs:166-184, content/src/segment.rs:171-183,292-360,
//              geometry/src/transform2d.rs:123-130,312-318
//   bin        renderer/src/tiler.rs:191-308, content/src/clip.rs:494-565,
//              renderer/src/builder.rs:509-616
//   propagate  renderer/src/tiler.rs:93-163 (backdrop prefix, clip cases), builder.rs:1013-1029 (z-buffer)
//   sort       shaders/d3d11/sort.cs.glsl:60-95 (painter's order + z-cull)
//   fill       shaders/fill_area.inc.glsl:11-27, shaders/d3d11/fill_compute.inc.glsl:11-25
//   tile       shaders/d3d11/tile.cs.glsl:71-163, shaders/tile_fragment.inc.glsl:539-614,
//              shaders/d3d9/tile_clip_combine.fs.glsl:28-31 (clip mask combine)
#include "kernels.cuh"

#include "common.cuh"
#include "scan.cuh"

namespace pf {

// SM count of the device current on this thread, cached per device (one process may drive several GPUs).
static int sm_count_of_current_device() {
    static int table[64] = {0};
    int dev = 0;
    PF_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) throw Error(2, "device ordinal out of range");
    if (table[dev] == 0) PF_CUDA_CHECK(cudaDeviceGetAttribute(&table[dev], cudaDevAttrMultiProcessorCount, dev));
    return table[dev];
}

// ---------------------------------------------------------------------------------------------
// Bit-exact scalar helpers (simd/src/x86/mod.rs semantics).
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float sse_min(float a, float b) { return a < b ? a : b; } // _mm_min_ps
__device__ __forceinline__ float sse_max(float a, float b) { return a > b ? a : b; } // _mm_max_ps
__device__ __forceinline__ int cvtps(float x) { return __float2int_rn(x); }          // _mm_cvtps_epi32

// Largest index p in [0, n) with keys[p] <= x, for non-decreasing keys with keys[0] <= x.
__device__ __forceinline__ uint32_t search_le(const uint32_t *__restrict__ keys, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) <= x)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// search_le through a coarse index: table[k] is the answer for x = k << shift, so the answer for
// any x lies in [table[x >> shift], table[(x >> shift) + 1]] — one or two probes instead of a
// 17-step dependent binary search over 100k+ paths.
__device__ __forceinline__ uint32_t search_coarse(const uint32_t *__restrict__ keys, const CoarseIndex ci, uint32_t x) {
    const uint32_t k = x >> ci.shift;
    uint32_t lo = __ldg(ci.table + k), hi = __ldg(ci.table + k + 1) + 1;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) <= x)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

__device__ __forceinline__ PathInfo load_path(const PathInfo *__restrict__ paths, uint32_t p) {
    const int4 *q = reinterpret_cast<const int4 *>(paths + p);
    int4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    PathInfo r;
    r.min_x = a.x, r.min_y = a.y, r.max_x = a.z, r.max_y = a.w;
    r.tile_offset = b.x, r.col_offset = b.y, r.seg_batch_first = b.z, r.seg_global_first = b.w;
    r.global_path_id = c.x, r.paint_ctrl = c.y, r.clip_path_index = c.z, r.pad = c.w;
    return r;
}

// ---------------------------------------------------------------------------------------------
// dice — Bézier flattening by recursive halving, one thread per input segment.
// ---------------------------------------------------------------------------------------------

struct Cubic {
    float2 p0, p1, p2, p3;
};

// Transform2F * Vector2F (geometry/src/transform2d.rs:123-130,312-318).
__device__ __forceinline__ float2 xf_apply(const Transform &t, float2 p) {
    if (t.identity) return p; // Outline::transform short-circuits (content/src/outline.rs:208-211)
    float hx = t.m11 * p.x, hy = t.m21 * p.x, hz = t.m12 * p.y, hw = t.m22 * p.y;
    return make_float2((hx + hz) + t.tx, (hy + hw) + t.ty);
}

// CubicSegment::is_flat(0.25) (content/src/segment.rs:292-300): flat iff the measure is <= 1.
__device__ __forceinline__ float cubic_flat_measure(const Cubic &c) {
    float u0x = ((3.0f * c.p1.x - c.p0.x) - c.p0.x) - c.p3.x;
    float u0y = ((3.0f * c.p1.y - c.p0.y) - c.p0.y) - c.p3.y;
    float u1x = ((3.0f * c.p2.x - c.p3.x) - c.p3.x) - c.p0.x;
    float u1y = ((3.0f * c.p2.y - c.p3.y) - c.p3.y) - c.p0.y;
    u0x = u0x * u0x, u0y = u0y * u0y, u1x = u1x * u1x, u1y = u1y * u1y;
    float mx = sse_max(u0x, u1x), my = sse_max(u0y, u1y);
    return mx + my; // compared with 16 * 0.25 * 0.25
}
__device__ __forceinline__ bool cubic_is_flat(const Cubic &c) { return cubic_flat_measure(c) <= 1.0f; }

__device__ __forceinline__ float2 lerp_half(float2 a, float2 b) { // a + t * (b - a), t = 0.5
    return make_float2(a.x + 0.5f * (b.x - a.x), a.y + 0.5f * (b.y - a.y));
}

constexpr int DICE_MAX_DEPTH = 40; // f32 halving collapses long before; the oracle uses the same cap
#ifndef PF_DICE_SMEM_LEVELS
#define PF_DICE_SMEM_LEVELS 5 // (8: 39 KB per block, 5 blocks per SM; 5: 26 KB, 7 blocks — random100k@8192 dice 0.210 -> 0.188 ms; deeper levels live in local memory)
#endif
constexpr int DICE_SMEM_LEVELS = PF_DICE_SMEM_LEVELS; // pending right halves kept in shared memory per thread
constexpr int DICE_THREADS = 128;

template <bool EMIT>
__global__ void __launch_bounds__(DICE_THREADS)
    k_dice(BatchDev b, uint32_t *__restrict__ seg_line_count, const uint32_t *__restrict__ seg_line_offset,
           float4 *__restrict__ lines, uint32_t *__restrict__ line_path, uint32_t line_capacity) {
    __shared__ float s_stack[DICE_SMEM_LEVELS][6][DICE_THREADS];
    __shared__ unsigned char s_depth[DICE_SMEM_LEVELS][DICE_THREADS];
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n_segments) return;
    uint32_t p = search_coarse(b.path_seg_first, b.seg_index, s);
    const PathInfo *pi = b.paths + p;
    uint32_t gseg = __ldg(&pi->seg_global_first) + (s - __ldg(&pi->seg_batch_first));
    uint2 si = __ldg(b.seg_indices + gseg);
    const float2 *pts = b.points + si.x;

    uint32_t out = EMIT ? seg_line_offset[s] : 0;
    uint32_t n = 0;
    auto emit_line = [&](float2 from, float2 to) {
        if (EMIT) {
            uint32_t o = out + n;
            if (o < line_capacity) {
                lines[o] = make_float4(from.x, from.y, to.x, to.y);
                line_path[o] = p;
            }
        }
        n++;
    };

    const bool is_cubic = (si.y & 0x40000000u) != 0, is_quad = (si.y & 0x80000000u) != 0;
    if (!is_cubic && !is_quad) {
        emit_line(xf_apply(b.xf, __ldg(pts)), xf_apply(b.xf, __ldg(pts + 1)));
    } else {
        Cubic cur;
        cur.p0 = xf_apply(b.xf, __ldg(pts));
        if (is_cubic) {
            cur.p1 = xf_apply(b.xf, __ldg(pts + 1));
            cur.p2 = xf_apply(b.xf, __ldg(pts + 2));
            cur.p3 = xf_apply(b.xf, __ldg(pts + 3));
        } else {
            // Segment::to_cubic (content/src/segment.rs:171-183)
            float2 c = xf_apply(b.xf, __ldg(pts + 1));
            cur.p3 = xf_apply(b.xf, __ldg(pts + 2));
            float2 c2 = make_float2(c.x + c.x, c.y + c.y);
            const float third = 1.0f / 3.0f;
            cur.p1 = make_float2((cur.p0.x + c2.x) * third, (cur.p0.y + c2.y) * third);
            cur.p2 = make_float2((c2.x + cur.p3.x) * third, (c2.y + cur.p3.y) * third);
        }
        // process_segment (renderer/src/tiler.rs:166-184) with an explicit stack: left half first.
        // A pending right half is stored as (p1, p2, p3, depth): its p0 is the end point of the
        // leaf emitted just before it is popped (splits keep end points bit-exact). The first
        // DICE_SMEM_LEVELS levels live in shared memory ([level][field][thread]: conflict-free);
        // deeper recursion, which f32 curves of sane size never reach, spills to local memory.
        float deep[DICE_MAX_DEPTH - DICE_SMEM_LEVELS][7];
        int sp = 0;
        int depth = 0;
        const int tid = threadIdx.x;
        for (;;) {
            if (cubic_is_flat(cur) || depth >= DICE_MAX_DEPTH) {
                emit_line(cur.p0, cur.p3);
                if (sp == 0) break;
                sp--;
                cur.p0 = cur.p3;
                if (sp < DICE_SMEM_LEVELS) {
                    cur.p1 = make_float2(s_stack[sp][0][tid], s_stack[sp][1][tid]);
                    cur.p2 = make_float2(s_stack[sp][2][tid], s_stack[sp][3][tid]);
                    cur.p3 = make_float2(s_stack[sp][4][tid], s_stack[sp][5][tid]);
                    depth = s_depth[sp][tid];
                } else {
                    const float *d = deep[sp - DICE_SMEM_LEVELS];
                    cur.p1 = make_float2(d[0], d[1]);
                    cur.p2 = make_float2(d[2], d[3]);
                    cur.p3 = make_float2(d[4], d[5]);
                    depth = (int)d[6];
                }
            } else {
                // CubicSegment::split(0.5) (content/src/segment.rs:307-360)
                float2 p01 = lerp_half(cur.p0, cur.p1), p12 = lerp_half(cur.p1, cur.p2),
                       p23 = lerp_half(cur.p2, cur.p3);
                float2 p012 = lerp_half(p01, p12), p123 = lerp_half(p12, p23);
                float2 p0123 = lerp_half(p012, p123);
                depth++;
                if (sp < DICE_SMEM_LEVELS) {
                    s_stack[sp][0][tid] = p123.x, s_stack[sp][1][tid] = p123.y;
                    s_stack[sp][2][tid] = p23.x, s_stack[sp][3][tid] = p23.y;
                    s_stack[sp][4][tid] = cur.p3.x, s_stack[sp][5][tid] = cur.p3.y;
                    s_depth[sp][tid] = (unsigned char)depth;
                } else {
                    float *d = deep[sp - DICE_SMEM_LEVELS];
                    d[0] = p123.x, d[1] = p123.y, d[2] = p23.x, d[3] = p23.y, d[4] = cur.p3.x, d[5] = cur.p3.y;
                    d[6] = (float)depth;
                }
                sp++;
                cur.p1 = p01, cur.p2 = p012, cur.p3 = p0123;
            }
        }
    }
    if (!EMIT) seg_line_count[s] = n;
}

// Single-pass dice for the steady state: lines are appended in arbitrary order, so no count pass and
// no scan are needed. Nothing downstream depends on the order of the lines — tile counts, backdrops
// and z values are accumulated with commutative atomics and coverage is summed as integers — only
// the parity dumps do, and they keep the ordered two-pass path above. The 32 lanes of a warp step
// their subdivision trees in lock-step; leaves produced in a step are compacted into a per-warp
// staging ring in shared memory and written out 32 at a time (one counter increment and one
// coalesced 512-byte store per 32 lines).
struct DiceShared {
    float stack[DICE_SMEM_LEVELS][8][DICE_THREADS]; // pending right halves, all four points (stealable)
    unsigned char depth[DICE_SMEM_LEVELS][DICE_THREADS];
    float4 line[DICE_THREADS / 32][64];
    uint32_t path[DICE_THREADS / 32][64];
};

// Loads and transforms segment s of the batch as a cubic; returns true for a straight line (p0 -> p3).
__device__ __forceinline__ bool load_segment(const BatchDev &b, uint32_t s, uint32_t &p, Cubic &cur) {
    p = search_coarse(b.path_seg_first, b.seg_index, s);
    const PathInfo *pi = b.paths + p;
    const uint32_t gseg = __ldg(&pi->seg_global_first) + (s - __ldg(&pi->seg_batch_first));
    const uint2 si = __ldg(b.seg_indices + gseg);
    const float2 *pts = b.points + si.x;
    const bool is_cubic = (si.y & 0x40000000u) != 0, is_quad = (si.y & 0x80000000u) != 0;
    cur.p0 = xf_apply(b.xf, __ldg(pts));
    if (is_cubic) {
        cur.p1 = xf_apply(b.xf, __ldg(pts + 1));
        cur.p2 = xf_apply(b.xf, __ldg(pts + 2));
        cur.p3 = xf_apply(b.xf, __ldg(pts + 3));
        return false;
    }
    if (is_quad) {
        // Segment::to_cubic (content/src/segment.rs:171-183)
        float2 c = xf_apply(b.xf, __ldg(pts + 1));
        cur.p3 = xf_apply(b.xf, __ldg(pts + 2));
        float2 c2 = make_float2(c.x + c.x, c.y + c.y);
        const float third = 1.0f / 3.0f;
        cur.p1 = make_float2((cur.p0.x + c2.x) * third, (cur.p0.y + c2.y) * third);
        cur.p2 = make_float2((c2.x + cur.p3.x) * third, (c2.y + cur.p3.y) * third);
        return false;
    }
    cur.p1 = cur.p2 = cur.p0;
    cur.p3 = xf_apply(b.xf, __ldg(pts + 1));
    return true;
}

// The lock-step subdivision of one (sub)curve per lane; must be called by all 32 lanes of the warp.
// Lanes whose curve is finished steal work: a lane's pending right halves form a deque in shared
// memory ([level][field][thread]); the owner pops from the top, an idle lane takes the bottom entry
// (the largest pending subtree) of a busy lane. Pairing is computed by every lane from two ballots,
// so no locks are needed, and a warp's step count approaches (nodes of its 32 curves) / 32 instead of
// the node count of its deepest curve.
__device__ __forceinline__ void dice_lockstep(DiceShared &sh, Cubic cur, int depth, bool active, bool is_line, uint32_t p,
                                              float4 *__restrict__ lines, uint32_t *__restrict__ line_path,
                                              uint32_t line_capacity, uint32_t *__restrict__ line_count) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lanes_below = (1u << lane) - 1u;
    float deep[DICE_MAX_DEPTH - DICE_SMEM_LEVELS][7];
    int sp = 0, bottom = 0; // this lane's pending right halves are levels [bottom, sp)
    // Staging ring of 64 lines per warp: [head, head + staged) mod 64 (both warp-uniform).
    uint32_t head = 0, staged = 0;
    auto flush = [&](uint32_t count) { // writes the oldest `count` (<= 32) staged lines
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(line_count, count);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t slot = (head + lane) & 63u;
        if ((uint32_t)lane < count && base + lane < line_capacity) {
            lines[base + lane] = sh.line[warp][slot];
            line_path[base + lane] = sh.path[warp][slot];
        }
        head = (head + count) & 63u;
        staged -= count;
    };
    while (__any_sync(0xffffffffu, active)) {
        bool leaf = false;
        float2 leaf_from = cur.p0, leaf_to = cur.p3;
        if (active) {
            // process_segment (renderer/src/tiler.rs:166-184), one node of the subdivision tree per step
            if (is_line || cubic_is_flat(cur) || depth >= DICE_MAX_DEPTH) {
                leaf = true;
                if (sp == bottom) {
                    active = false;
                } else {
                    sp--;
                    cur.p0 = cur.p3; // = the stored p0: splits keep end points bit-exact
                    if (sp < DICE_SMEM_LEVELS) {
                        cur.p1 = make_float2(sh.stack[sp][2][tid], sh.stack[sp][3][tid]);
                        cur.p2 = make_float2(sh.stack[sp][4][tid], sh.stack[sp][5][tid]);
                        cur.p3 = make_float2(sh.stack[sp][6][tid], sh.stack[sp][7][tid]);
                        depth = sh.depth[sp][tid];
                    } else {
                        const float *d = deep[sp - DICE_SMEM_LEVELS];
                        cur.p1 = make_float2(d[0], d[1]);
                        cur.p2 = make_float2(d[2], d[3]);
                        cur.p3 = make_float2(d[4], d[5]);
                        depth = (int)d[6];
                    }
                }
            } else {
                // CubicSegment::split(0.5) (content/src/segment.rs:307-360)
                float2 p01 = lerp_half(cur.p0, cur.p1), p12 = lerp_half(cur.p1, cur.p2),
                       p23 = lerp_half(cur.p2, cur.p3);
                float2 p012 = lerp_half(p01, p12), p123 = lerp_half(p12, p23);
                float2 p0123 = lerp_half(p012, p123);
                depth++;
                if (sp < DICE_SMEM_LEVELS) {
                    sh.stack[sp][0][tid] = p0123.x, sh.stack[sp][1][tid] = p0123.y;
                    sh.stack[sp][2][tid] = p123.x, sh.stack[sp][3][tid] = p123.y;
                    sh.stack[sp][4][tid] = p23.x, sh.stack[sp][5][tid] = p23.y;
                    sh.stack[sp][6][tid] = cur.p3.x, sh.stack[sp][7][tid] = cur.p3.y;
                    sh.depth[sp][tid] = (unsigned char)depth;
                } else {
                    float *d = deep[sp - DICE_SMEM_LEVELS];
                    d[0] = p123.x, d[1] = p123.y, d[2] = p23.x, d[3] = p23.y, d[4] = cur.p3.x, d[5] = cur.p3.y;
                    d[6] = (float)depth;
                }
                sp++;
                cur.p1 = p01, cur.p2 = p012, cur.p3 = p0123;
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, leaf);
        if (leaf) {
            const uint32_t slot = (head + staged + __popc(ballot & lanes_below)) & 63u;
            sh.line[warp][slot] = make_float4(leaf_from.x, leaf_from.y, leaf_to.x, leaf_to.y);
            sh.path[warp][slot] = p;
        }
        staged += __popc(ballot);
        if (staged >= 32) {
            __syncwarp();
            flush(32);
            __syncwarp(); // the flushed slots may be rewritten by the next step
        }
        // ---- work stealing: the k-th idle lane takes the bottom entry of the k-th lane that can give one
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0) continue;
        const bool can_give = active && bottom < sp && bottom < DICE_SMEM_LEVELS;
        const uint32_t givers = __ballot_sync(0xffffffffu, can_give);
        if (givers == 0) continue;
        __syncwarp(); // the givers' pushes of this step are visible
        const int pairs = min(__popc(idle), __popc(givers));
        int victim = -1;
        if (!active && __popc(idle & lanes_below) < pairs) victim = (int)__fns(givers, 0, __popc(idle & lanes_below) + 1);
        const int victim_bottom = __shfl_sync(0xffffffffu, bottom, victim < 0 ? 0 : victim);
        const uint32_t victim_path = __shfl_sync(0xffffffffu, p, victim < 0 ? 0 : victim);
        if (victim >= 0) {
            const int vt = (tid & ~31) + victim;
            cur.p0 = make_float2(sh.stack[victim_bottom][0][vt], sh.stack[victim_bottom][1][vt]);
            cur.p1 = make_float2(sh.stack[victim_bottom][2][vt], sh.stack[victim_bottom][3][vt]);
            cur.p2 = make_float2(sh.stack[victim_bottom][4][vt], sh.stack[victim_bottom][5][vt]);
            cur.p3 = make_float2(sh.stack[victim_bottom][6][vt], sh.stack[victim_bottom][7][vt]);
            depth = sh.depth[victim_bottom][vt];
            p = victim_path;
            is_line = false;
            active = true;
            sp = bottom = 0;
        } else if (can_give && __popc(givers & lanes_below) < pairs) {
            bottom++;
        }
        __syncwarp(); // stolen entries are read before their owners can overwrite anything
    }
    __syncwarp();
    if (staged) flush(staged);
    __syncwarp();
}

__global__ void __launch_bounds__(DICE_THREADS)
    k_dice_stream(BatchDev b, uint32_t segments_per_warp, float4 *__restrict__ lines, uint32_t *__restrict__ line_path,
                  uint32_t line_capacity, uint32_t *__restrict__ line_count) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    __shared__ DiceShared sh;
    // A warp starts with segments_per_warp (<= 32) curves, one per low lane: on small scenes the other
    // lanes start idle and take the first right halves that appear, so even a few thousand curves
    // spread over every SM and the deepest curve is shared by a whole warp.
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t s = warp_global * segments_per_warp + lane;
    bool active = lane < segments_per_warp && s < b.n_segments;
    bool is_line = false;
    uint32_t p = 0;
    Cubic cur;
    cur.p0 = cur.p1 = cur.p2 = cur.p3 = make_float2(0.0f, 0.0f);
    if (active) is_line = load_segment(b, s, p, cur);
    dice_lockstep(sh, cur, 0, active, is_line, p, lines, line_path, line_capacity, line_count);
}

int launch_dice_stream(const BatchDev &b, float4 *lines, uint32_t *line_path, uint32_t line_capacity,
                       uint32_t *line_count, cudaStream_t stream) {
    if (b.n_segments == 0) return 0;
    // Enough warps to give every scheduler of the GPU a few (148 SMs x 4 schedulers x 2), at most 32 curves each.
    const int sm_count = sm_count_of_current_device();
    const uint32_t target_warps = (uint32_t)sm_count * 8u;
    const uint32_t per_warp = std::min(32u, std::max(1u, div_up(b.n_segments, target_warps)));
    const uint32_t warps = div_up(b.n_segments, per_warp);
    launch_chained(k_dice_stream, div_up(warps, DICE_THREADS / 32), DICE_THREADS, stream, b, per_warp, lines, line_path,
                   line_capacity, line_count);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

int launch_dice(bool emit, const BatchDev &b, uint32_t *seg_line_count, const uint32_t *seg_line_offset,
                float4 *lines, uint32_t *line_path, uint32_t line_capacity, cudaStream_t stream) {
    if (b.n_segments == 0) return 0;
    unsigned grid = div_up(b.n_segments, DICE_THREADS);
    if (emit)
        k_dice<true><<<grid, DICE_THREADS, 0, stream>>>(b, seg_line_count, seg_line_offset, lines, line_path, line_capacity);
    else
        k_dice<false><<<grid, DICE_THREADS, 0, stream>>>(b, seg_line_count, seg_line_offset, lines, line_path, line_capacity);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// ---------------------------------------------------------------------------------------------
// bin — segment-to-tile lattice clipping, one thread per flattened line.
//
// Three modes of one walk:
//   BIN_COUNT      per-tile fill counts + backdrop deltas (+ fills per line for the parity dumps)
//   BIN_EMIT_LIVE  production emit, run after propagate + z-cull: only the fills of tiles that
//                  survived are stored, straight into their tile-grouped runs
//   BIN_EMIT       parity dumps: every fill, tile-grouped and at its emission-order slot (what
//                  AddFillsD3D9 carries)
// Lines that stay inside one tile (the majority after flattening) take a division-free fast path;
// the others are compacted through shared memory so the long walks run on dense warps.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + (b - a) * t; } // util.rs:25-27

// clip_line_segment_to_rect (content/src/clip.rs:494-565) against
// [vb.min_x, vb.max_x] x [-inf, vb.max_y] (renderer/src/tiler.rs:194-200).
__device__ __forceinline__ bool clip_line(float2 &from, float2 &to, const ViewBox &vb) {
    auto outcode = [&](float2 q) -> unsigned {
        unsigned o = 0;
        if (q.x < vb.min_x) o |= 1u; // LEFT
        if (q.x > vb.max_x) o |= 2u; // RIGHT
        if (q.y > vb.max_y) o |= 8u; // BOTTOM   (TOP is -inf: never set)
        return o;
    };
    unsigned of = outcode(from), ot = outcode(to);
    for (;;) {
        if ((of | ot) == 0) return true;
        if ((of & ot) != 0) return false;
        bool clip_from = of > ot;
        float2 point = clip_from ? from : to;
        unsigned oc = clip_from ? of : ot;
        if (oc & 1u) {
            point = make_float2(vb.min_x, lerpf(from.y, to.y, (vb.min_x - from.x) / (to.x - from.x)));
        } else if (oc & 2u) {
            point = make_float2(vb.max_x, lerpf(from.y, to.y, (vb.max_x - from.x) / (to.x - from.x)));
        } else if (oc & 8u) {
            point = make_float2(lerpf(from.x, to.x, (vb.max_y - from.y) / (to.y - from.y)), vb.max_y);
        }
        if (clip_from) {
            from = point;
            of = outcode(point);
        } else {
            to = point;
            ot = outcode(point);
        }
    }
}

enum BinMode { BIN_EMIT_LIVE = 0, BIN_COUNT = 1, BIN_EMIT = 2 };
constexpr int BIN_THREADS = 128;

// Receives the fills and backdrop adjustments of one line.
template <int MODE>
struct BinSink {
    const BinArgs &a;
    const PathInfo &path;
    int rect_w, rect_h;
    uint32_t out_base;
    uint32_t emitted;

    // ObjectBuilder::add_fill (renderer/src/builder.rs:509-553)
    __device__ __forceinline__ void add_fill(float2 from, float2 to, int tx, int ty) {
        int ox = tx - path.min_x, oy = ty - path.min_y;
        if (ox < 0 || oy < 0 || ox >= rect_w || oy >= rect_h) return; // tile_coords_to_local_index
        float ulx = (float)tx * 16.0f, uly = (float)ty * 16.0f;
        int fx = cvtps(sse_min(sse_max((from.x - ulx) * 256.0f, 0.0f), 4095.0f));
        int fy = cvtps(sse_min(sse_max((from.y - uly) * 256.0f, 0.0f), 4095.0f));
        int tx8 = cvtps(sse_min(sse_max((to.x - ulx) * 256.0f, 0.0f), 4095.0f));
        int ty8 = cvtps(sse_min(sse_max((to.y - uly) * 256.0f, 0.0f), 4095.0f));
        if (fx == tx8) return; // cull degenerate fills
        const uint32_t t = path.tile_offset + (uint32_t)(ox + rect_w * oy);
        const uint32_t from_w = (uint32_t)fx | ((uint32_t)fy << 16), to_w = (uint32_t)tx8 | ((uint32_t)ty8 << 16);
        if (MODE == BIN_EMIT_LIVE) {
            // Production emit, after the z-cull: tiles that lost the z-test (or lie outside the
            // framebuffer) get no space in the tile-grouped array — occlusion culling before fill
            // emission.
            if (__ldg(a.tile_fb + t) != 0xffffffffu) {
                uint32_t pos = atomicAdd(a.tile_fill_pos + t, 1u);
                if (pos < a.fill_capacity) a.fills[pos] = make_uint2(from_w, to_w);
            }
        } else if (MODE == BIN_COUNT) {
            atomicAdd(a.tile_word + t, 1u);
        } else {
            // Parity dumps: every fill tile-grouped (no culling) and at its emission-order slot.
            uint32_t pos = atomicAdd(a.tile_fill_pos + t, 1u);
            if (pos < a.fill_capacity) a.fills[pos] = make_uint2(from_w, to_w);
            uint32_t e = out_base + emitted;
            atomicMin(a.tile_first_fill + t, e);
            if (e < a.emit_capacity) a.fills_emit[e] = EmitFill{from_w, to_w, t};
        }
        emitted++;
    }

    // ObjectBuilder::adjust_alpha_tile_backdrop (renderer/src/builder.rs:595-616); once per frame.
    __device__ __forceinline__ void adjust_backdrop(int tx, int ty, int delta) {
        if (MODE != BIN_COUNT) return;
        int ox = tx - path.min_x, oy = ty - path.min_y;
        if (ox < 0 || ox >= rect_w || oy >= rect_h) return;
        if (oy < 0) {
            atomicAdd(a.col_backdrop + path.col_offset + ox, delta);
            return;
        }
        // i8 wrapping add in the top byte of the tile word (no carry reaches the count bits).
        atomicAdd(a.tile_word + path.tile_offset + (uint32_t)(ox + rect_w * oy), (uint32_t)delta << 24);
    }
};

// process_line_segment (renderer/src/tiler.rs:202-308) after clipping, for a line that crosses at
// least one tile boundary, split into "advance the Amanatides-Woo state by one tile" (cheap, strictly
// serial: t_max is accumulated) and "emit the fills / backdrop change of that tile" (the expensive
// part, independent per tile).
struct WalkState {
    float2 from;
    float vx, vy, t_max_x, t_max_y, t_delta_x, t_delta_y;
    float2 cur;
    int step_x, step_y, tx, ty, to_tx, to_ty, last_step; // step codes: 0 none, 1 X, 2 Y
};
struct WalkStep {
    float2 cur, next;
    int tx, ty, last_step, next_step;
};

__device__ __forceinline__ void walk_init(WalkState &w, float2 from, float2 to, int from_tx, int from_ty, int to_tx,
                                          int to_ty) {
    const float tile_size = 16.0f;
    w.from = from;
    w.vx = to.x - from.x, w.vy = to.y - from.y;
    const bool neg_x = w.vx < 0.0f, neg_y = w.vy < 0.0f;
    w.step_x = neg_x ? -1 : 1, w.step_y = neg_y ? -1 : 1;
    const float first_cross_x = (float)(from_tx + (neg_x ? 0 : 1)) * tile_size;
    const float first_cross_y = (float)(from_ty + (neg_y ? 0 : 1)) * tile_size;
    w.t_max_x = (first_cross_x - from.x) / w.vx;
    w.t_max_y = (first_cross_y - from.y) / w.vy;
    w.t_delta_x = fabsf(tile_size / w.vx);
    w.t_delta_y = fabsf(tile_size / w.vy);
    w.cur = from;
    w.tx = from_tx, w.ty = from_ty, w.to_tx = to_tx, w.to_ty = to_ty;
    w.last_step = 0;
}

// Fills `st` with the current tile's step and moves to the next tile. Returns false after the last tile.
__device__ __forceinline__ bool walk_advance(WalkState &w, WalkStep &st) {
    int next_step;
    if (w.t_max_x < w.t_max_y)
        next_step = 1;
    else if (w.t_max_x > w.t_max_y)
        next_step = 2;
    else
        next_step = w.step_x > 0 ? 1 : 2;
    const float next_t = fminf(next_step == 1 ? w.t_max_x : w.t_max_y, 1.0f);
    if (w.tx == w.to_tx && w.ty == w.to_ty) next_step = 0;
    const float2 next = make_float2(w.from.x + w.vx * next_t, w.from.y + w.vy * next_t);
    st.cur = w.cur, st.next = next, st.tx = w.tx, st.ty = w.ty, st.last_step = w.last_step, st.next_step = next_step;
    if (next_step == 0) return false;
    if (next_step == 1) {
        if (w.tx == w.to_tx) return false;
        w.t_max_x += w.t_delta_x;
        w.t_max_y += 0.0f;
        w.tx += w.step_x;
    } else {
        if (w.ty == w.to_ty) return false;
        w.t_max_x += 0.0f;
        w.t_max_y += w.t_delta_y;
        w.ty += w.step_y;
    }
    w.cur = next;
    w.last_step = next_step;
    return true;
}

template <typename Sink>
__device__ __forceinline__ void walk_emit(const WalkStep &st, int step_x, int step_y, Sink &sink) {
    const float tile_size = 16.0f;
    sink.add_fill(st.cur, st.next, st.tx, st.ty);
    if (step_y < 0 && st.next_step == 2) { // leaves through the top boundary
        sink.add_fill(st.next, make_float2((float)st.tx * tile_size, (float)st.ty * tile_size), st.tx, st.ty);
    } else if (step_y > 0 && st.last_step == 2) { // entered through the top boundary
        sink.add_fill(make_float2((float)st.tx * tile_size, (float)st.ty * tile_size), st.cur, st.tx, st.ty);
    }
    if (step_x < 0 && st.last_step == 1) {
        sink.adjust_backdrop(st.tx, st.ty, 1); // entered through the right boundary
    } else if (step_x > 0 && st.next_step == 1) {
        sink.adjust_backdrop(st.tx, st.ty, -1); // leaving through the right boundary
    }
}

// The walk by one thread (the common case), written out so the state stays in registers.
template <typename Sink>
__device__ __forceinline__ void walk_line(float2 from, float2 to, int from_tx, int from_ty, int to_tx, int to_ty,
                                          Sink &sink) {
    const float tile_size = 16.0f;
    float vx = to.x - from.x, vy = to.y - from.y;
    bool neg_x = vx < 0.0f, neg_y = vy < 0.0f;
    int step_x = neg_x ? -1 : 1, step_y = neg_y ? -1 : 1;
    float first_cross_x = (float)(from_tx + (neg_x ? 0 : 1)) * tile_size;
    float first_cross_y = (float)(from_ty + (neg_y ? 0 : 1)) * tile_size;
    float t_max_x = (first_cross_x - from.x) / vx;
    float t_max_y = (first_cross_y - from.y) / vy;
    float t_delta_x = fabsf(tile_size / vx);
    float t_delta_y = fabsf(tile_size / vy);

    float2 cur = from;
    int tx = from_tx, ty = from_ty;
    int last_step = 0; // 0 none, 1 X, 2 Y
    for (;;) {
        int next_step;
        if (t_max_x < t_max_y)
            next_step = 1;
        else if (t_max_x > t_max_y)
            next_step = 2;
        else
            next_step = step_x > 0 ? 1 : 2;
        float next_t = fminf(next_step == 1 ? t_max_x : t_max_y, 1.0f);
        if (tx == to_tx && ty == to_ty) next_step = 0;
        float2 next = make_float2(from.x + vx * next_t, from.y + vy * next_t);
        sink.add_fill(cur, next, tx, ty);
        if (step_y < 0 && next_step == 2) {
            sink.add_fill(next, make_float2((float)tx * tile_size, (float)ty * tile_size), tx, ty);
        } else if (step_y > 0 && last_step == 2) {
            sink.add_fill(make_float2((float)tx * tile_size, (float)ty * tile_size), cur, tx, ty);
        }
        if (step_x < 0 && last_step == 1) {
            sink.adjust_backdrop(tx, ty, 1);
        } else if (step_x > 0 && next_step == 1) {
            sink.adjust_backdrop(tx, ty, -1);
        }
        if (next_step == 0) break;
        if (next_step == 1) {
            if (tx == to_tx) break;
            t_max_x += t_delta_x;
            t_max_y += 0.0f;
            tx += step_x;
        } else {
            if (ty == to_ty) break;
            t_max_x += 0.0f;
            t_max_y += t_delta_y;
            ty += step_y;
        }
        cur = next;
        last_step = next_step;
    }
}

// The same walk by a whole warp, for lines that cross many tiles (long straight edges). The walk is serial — t_max is
// accumulated one rounding at a time — but what is loop-carried is small: the two t_max, the tile and the step taken.
// The walk visits at most |dx| + |dy| + 1 tiles; lane k takes the k-th chunk of ceil(that / 32) consecutive tiles. It
// first replays the steps before its chunk keeping only the loop-carried state (a compare, a select and one add per
// step: the same operations on the same values as the serial walk, so the same bits), then walks its own tiles in full:
// end points of a step (from + v * t, the expression the serial walk evaluates, on the same t), fills, backdrop change.
// A 256-tile edge costs 31 * 8 light steps + 8 full ones instead of 256 full ones. Must be called by all 32 lanes with
// identical arguments. Fill order within the line is not preserved (only BIN_EMIT needs it).
template <typename Sink>
__device__ __forceinline__ void walk_line_warp(float2 from, float2 to, int from_tx, int from_ty, int to_tx, int to_ty,
                                               Sink &sink) {
    const int lane = threadIdx.x & 31;
    WalkState w;
    walk_init(w, from, to, from_tx, from_ty, to_tx, to_ty);
    float t_max_x = w.t_max_x, t_max_y = w.t_max_y;
    int tx = from_tx, ty = from_ty, last_step = 0;
    float prev_t = 0.0f;
    bool first = true, more = true;
    const int chunk = (abs(to_tx - from_tx) + abs(to_ty - from_ty) + 1 + 31) >> 5;
    const bool tie_x = w.step_x > 0;
    for (int k = lane * chunk; k > 0 && more; k--) {
        const bool step_in_x = t_max_x < t_max_y || (!(t_max_x > t_max_y) && tie_x);
        if (tx == to_tx && ty == to_ty) { // the walk ends before this lane's chunk
            more = false;
            break;
        }
        const float chosen = step_in_x ? t_max_x : t_max_y;
        if (step_in_x) {
            if (tx == to_tx) more = false;
            t_max_x += w.t_delta_x;
            tx += w.step_x;
        } else {
            if (ty == to_ty) more = false;
            t_max_y += w.t_delta_y;
            ty += w.step_y;
        }
        prev_t = fminf(chosen, 1.0f);
        last_step = step_in_x ? 1 : 2;
        first = false;
    }
    for (int c = 0; c < chunk && more; c++) {
        int next_step;
        if (t_max_x < t_max_y)
            next_step = 1;
        else if (t_max_x > t_max_y)
            next_step = 2;
        else
            next_step = tie_x ? 1 : 2;
        const float next_t = fminf(next_step == 1 ? t_max_x : t_max_y, 1.0f);
        if (tx == to_tx && ty == to_ty) next_step = 0;
        WalkStep st;
        st.cur = first ? from : make_float2(from.x + w.vx * prev_t, from.y + w.vy * prev_t);
        st.next = make_float2(from.x + w.vx * next_t, from.y + w.vy * next_t);
        st.tx = tx, st.ty = ty, st.last_step = last_step, st.next_step = next_step;
        walk_emit(st, w.step_x, w.step_y, sink);
        if (next_step == 0) {
            more = false;
        } else if (next_step == 1) {
            if (tx == to_tx) more = false;
            t_max_x += w.t_delta_x;
            tx += w.step_x;
        } else {
            if (ty == to_ty) more = false;
            t_max_y += w.t_delta_y;
            ty += w.step_y;
        }
        prev_t = next_t;
        last_step = next_step;
        first = false;
    }
    __syncwarp();
}

constexpr int BIN_LONG_STEPS = 12; // tile crossings from which a line is walked by a whole warp

#ifndef PF_BIN_MIN_BLOCKS
#define PF_BIN_MIN_BLOCKS 16 // 32 registers: all 64 warps of an SM resident (latency-bound: random100k@8192 bin 0.250 -> 0.228 ms)
#endif
template <int MODE>
__global__ void __launch_bounds__(BIN_THREADS, PF_BIN_MIN_BLOCKS) k_bin(BatchDev b, BinArgs a) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    __shared__ float4 s_line[BIN_THREADS];
    __shared__ uint32_t s_index[BIN_THREADS];
    __shared__ uint32_t s_queued;
    if (threadIdx.x == 0) s_queued = 0;
    __syncthreads();

    const uint32_t n_lines = a.n_lines_dev ? min(a.n_lines, __ldg(a.n_lines_dev)) : a.n_lines;
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    const float recip = 1.0f / 16.0f;
    if (l < n_lines && (MODE != BIN_EMIT_LIVE || __ldg(a.path_live + __ldg(a.line_path + l)) != 0u)) {
        // (BIN_EMIT_LIVE: a path whose tiles were all culled has nothing left to emit)
        float4 seg = __ldg(a.lines + l);
        const PathInfo path = load_path(b.paths, __ldg(a.line_path + l));
        const int rect_w = path.max_x - path.min_x, rect_h = path.max_y - path.min_y;
        float2 from = make_float2(seg.x, seg.y), to = make_float2(seg.z, seg.w);
        uint32_t emitted = 0;
        if (rect_w > 0 && clip_line(from, to, b.view_box)) {
            int from_tx = cvtps(floorf(from.x * recip)), from_ty = cvtps(floorf(from.y * recip));
            int to_tx = cvtps(floorf(to.x * recip)), to_ty = cvtps(floorf(to.y * recip));
            if (from_tx == to_tx && from_ty == to_ty) {
                // The line stays inside one tile: both t_max are >= 1 (the first crossing lies
                // beyond `to`), so next_t = min(t_max, 1) = 1, there is no step, no auxiliary fill
                // and no backdrop change — one fill from `from` to from + v * 1.
                float vx = to.x - from.x, vy = to.y - from.y;
                BinSink<MODE> sink{a, path, rect_w, rect_h,
                                   (MODE == BIN_EMIT) ? __ldg(a.line_fill_offset + l) : 0u, 0u};
                sink.add_fill(from, make_float2(from.x + vx * 1.0f, from.y + vy * 1.0f), from_tx, from_ty);
                emitted = sink.emitted;
            } else {
                uint32_t q = atomicAdd(&s_queued, 1u);
                s_line[q] = make_float4(from.x, from.y, to.x, to.y);
                s_index[q] = l;
                emitted = 0xffffffffu; // counted by whichever thread walks it
            }
        }
        if (MODE == BIN_COUNT && a.line_fill_count && emitted != 0xffffffffu) a.line_fill_count[l] = emitted;
    }
    __syncthreads();
    // Long walks, compacted: thread q takes the q-th queued line. Lines that cross many tiles go to
    // a global queue and are walked by whole warps in k_bin_long (kept out of this kernel so its
    // register count, and with it the occupancy of the common path, stays low).
    const uint32_t queued = s_queued;
    for (uint32_t q = threadIdx.x; q < queued; q += BIN_THREADS) {
        const float4 seg = s_line[q];
        const uint32_t li = s_index[q];
        float2 from = make_float2(seg.x, seg.y), to = make_float2(seg.z, seg.w);
        int from_tx = cvtps(floorf(from.x * recip)), from_ty = cvtps(floorf(from.y * recip));
        int to_tx = cvtps(floorf(to.x * recip)), to_ty = cvtps(floorf(to.y * recip));
        if (MODE != BIN_EMIT && a.long_queue && abs(to_tx - from_tx) + abs(to_ty - from_ty) >= BIN_LONG_STEPS) {
            const uint32_t slot = atomicAdd(a.long_count, 1u);
            if (slot < a.long_capacity) {
                a.long_queue[slot] = li;
                continue;
            }
        }
        const PathInfo path = load_path(b.paths, __ldg(a.line_path + li));
        const int rect_w = path.max_x - path.min_x, rect_h = path.max_y - path.min_y;
        BinSink<MODE> sink{a, path, rect_w, rect_h, (MODE == BIN_EMIT) ? __ldg(a.line_fill_offset + li) : 0u, 0u};
        walk_line(from, to, from_tx, from_ty, to_tx, to_ty, sink);
        if (MODE == BIN_COUNT && a.line_fill_count) a.line_fill_count[li] = sink.emitted;
    }
}

// Lines that cross >= BIN_LONG_STEPS tiles (long straight edges: the tiger at 4K has edges spanning
// hundreds of tiles), one warp per line, warps pull from the queue the main kernel filled.
template <int MODE>
__global__ void __launch_bounds__(BIN_THREADS) k_bin_long(BatchDev b, BinArgs a) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    const uint32_t n_long = min(*a.long_count, a.long_capacity);
    const float recip = 1.0f / 16.0f;
    for (;;) {
        uint32_t k = 0;
        if ((threadIdx.x & 31) == 0) k = atomicAdd(a.long_cursor, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_long) return;
        const uint32_t li = a.long_queue[k];
        const float4 seg = __ldg(a.lines + li);
        const PathInfo path = load_path(b.paths, __ldg(a.line_path + li));
        const int rect_w = path.max_x - path.min_x, rect_h = path.max_y - path.min_y;
        float2 from = make_float2(seg.x, seg.y), to = make_float2(seg.z, seg.w);
        clip_line(from, to, b.view_box); // it was accepted by the main kernel: same result
        int from_tx = cvtps(floorf(from.x * recip)), from_ty = cvtps(floorf(from.y * recip));
        int to_tx = cvtps(floorf(to.x * recip)), to_ty = cvtps(floorf(to.y * recip));
        BinSink<MODE> sink{a, path, rect_w, rect_h, 0u, 0u};
        walk_line_warp(from, to, from_tx, from_ty, to_tx, to_ty, sink);
        if (MODE == BIN_COUNT && a.line_fill_count) {
            uint32_t total = sink.emitted;
            for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
            if ((threadIdx.x & 31) == 0) a.line_fill_count[li] = total;
        }
    }
}

int launch_bin(int mode, const BatchDev &b, const BinArgs &args, cudaStream_t stream) {
    if (args.n_lines == 0) return 0;
    unsigned grid = div_up(args.n_lines, BIN_THREADS);
    int launches = 1;
    // (the caller hands every pass its own zeroed long_count / long_cursor words)
    if (mode == BIN_EMIT_LIVE)
        launch_chained(k_bin<BIN_EMIT_LIVE>, grid, BIN_THREADS, stream, b, args);
    else if (mode == BIN_COUNT)
        launch_chained(k_bin<BIN_COUNT>, grid, BIN_THREADS, stream, b, args);
    else
        k_bin<BIN_EMIT><<<grid, BIN_THREADS, 0, stream>>>(b, args);
    if (mode != BIN_EMIT && args.long_queue) {
        const unsigned long_grid = (unsigned)sm_count_of_current_device() * 2u; // persistent warps; exits at once when the queue is empty
        if (mode == BIN_EMIT_LIVE)
            launch_chained(k_bin_long<BIN_EMIT_LIVE>, long_grid, BIN_THREADS, stream, b, args);
        else
            launch_chained(k_bin_long<BIN_COUNT>, long_grid, BIN_THREADS, stream, b, args);
        launches++;
    }
    PF_CUDA_CHECK(cudaGetLastError());
    return launches;
}

// Sum of the per-tile fill counts (RenderStats.fill_count), computed on demand.
__global__ void __launch_bounds__(256) k_sum_fill_counts(const uint32_t *__restrict__ tile_word, uint32_t n_tiles,
                                                         unsigned long long *__restrict__ total) {
    unsigned long long sum = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles; i += (size_t)gridDim.x * blockDim.x)
        sum += tile_word[i] & 0x00ffffffu;
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(total, sum);
}
int launch_sum_fill_counts(const uint32_t *tile_word, uint32_t n_tiles, unsigned long long *total, cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    unsigned grid = div_up(n_tiles, 256 * 8);
    k_sum_fill_counts<<<grid < 1184 ? grid : 1184, 256, 0, stream>>>(tile_word, n_tiles, total);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// ---------------------------------------------------------------------------------------------
// propagate — backdrop prefix sums down tile columns + occluder z-writes, one thread per column.
// ---------------------------------------------------------------------------------------------

// Two register budgets: a batch with enough columns to fill the GPU runs at 32 registers with every warp slot of an
// SM taken (random100k@8192: 1.05 M columns, 0.053 -> 0.048 ms); a batch of few, tall columns (tiger@4096: 9.6 k columns
// of up to 256 rows) is one serial chain per thread whatever the occupancy, and the tighter budget only costs it
// (0.033 -> 0.040 ms), so it keeps the registers the compiler asks for.
constexpr uint32_t PROPAGATE_DENSE_COLUMNS = 262144;
template <bool HAS_CLIP, bool DENSE>
__global__ void __launch_bounds__(128, DENSE ? 16 : 8)
    k_propagate(BatchDev b, uint32_t *__restrict__ tile_word, const int32_t *__restrict__ col_backdrop,
                int32_t *__restrict__ z_buffer, ClipDev clip, uint32_t *__restrict__ tile_clip,
                uint32_t *__restrict__ tile_orig_count) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.n_columns) return;
    uint32_t p = search_coarse(b.path_col_offset, b.col_index, c);
    const PathInfo path = load_path(b.paths, p);
    const int w = path.max_x - path.min_x, h = path.max_y - path.min_y;
    const int x = (int)(c - path.col_offset);
    const bool z_write = (path.paint_ctrl >> 24) & 1u;
    const int fb_w = b.fb.max_x - b.fb.min_x;
    const int fx = path.min_x + x - b.fb.min_x;
    const bool fx_ok = fx >= 0 && fx < fb_w;
    int32_t backdrop = col_backdrop[c];
    uint32_t t = path.tile_offset + (uint32_t)x;
    const int fb_h = b.fb.max_y - b.fb.min_y;
    // The clip path of this draw path, if any: the column's clip tiles are at clip_t0 + row * clip_w.
    bool clipped = false, clip_column = false;
    int clip_w = 0, clip_min_y = 0, clip_max_y = 0;
    uint32_t clip_t0 = 0;
    if (HAS_CLIP && path.clip_path_index < clip.n_paths) {
        clipped = true;
        const PathInfo cp = load_path(clip.paths, path.clip_path_index);
        const int cx = path.min_x + x - cp.min_x;
        clip_w = cp.max_x - cp.min_x, clip_min_y = cp.min_y, clip_max_y = cp.max_y;
        clip_column = cx >= 0 && cx < clip_w;
        clip_t0 = cp.tile_offset + (uint32_t)cx;
    }
    // The prefix sum is serial down the column, but the loads are not: fetch 8 rows ahead so the
    // chain waits for memory once per 8 tiles instead of once per tile.
    constexpr int AHEAD = 8;
    for (int y0 = 0; y0 < h; y0 += AHEAD) {
        uint32_t words[AHEAD];
#pragma unroll
        for (int k = 0; k < AHEAD; k++)
            words[k] = (y0 + k < h) ? __ldcg(tile_word + t + (uint32_t)(k * w)) : 0u;
#pragma unroll
        for (int k = 0; k < AHEAD; k++) {
            const int y = y0 + k;
            if (y < h) {
                const uint32_t word = words[k];
                const int delta = (int)(int8_t)(word >> 24);
                uint32_t count = word & 0x00ffffffu;
                int8_t b8 = (int8_t)backdrop; // backdrops[column] as i8   (renderer/src/tiler.rs:112)
                uint32_t clip_ref = 0;
                // (parity dumps: the fills of a tile the clip drops stay in the fill list with their alpha tile id)
                if (HAS_CLIP && tile_orig_count) tile_orig_count[t + (uint32_t)(k * w)] = count;
                if (HAS_CLIP && !clipped) tile_clip[t + (uint32_t)(k * w)] = 0;
                if (HAS_CLIP && clipped) {
                    // Tiler::prepare_tiles, the four clip cases (renderer/src/tiler.rs:114-156;
                    // twin: shaders/d3d11/propagate.cs.glsl:142-189).
                    const int ty = path.min_y + y;
                    bool drop = true;
                    if (clip_column && ty >= clip_min_y && ty < clip_max_y) {
                        const uint32_t ct = clip_t0 + (uint32_t)(ty - clip_min_y) * (uint32_t)clip_w;
                        const uint32_t cw = __ldg(clip.tile_word + ct);
                        const bool clip_alpha = (cw & 0x00ffffffu) != 0;
                        drop = false;
                        if (clip_alpha && count != 0) {
                            clip_ref = ct + 1u; // both are masks: min-combined in the fused kernel
                        } else if (clip_alpha && count == 0 && b8 != 0) {
                            clip_ref = (ct + 1u) | TILE_CLIP_REPLACE; // the solid draw tile takes the clip tile's mask
                        } else if (!clip_alpha && (cw >> 24) == 0) {
                            drop = true; // the clip path does not cover this tile
                        }
                    }
                    if (drop) count = 0, b8 = 0;
                    tile_clip[t + (uint32_t)(k * w)] = clip_ref;
                }
                tile_word[t + (uint32_t)(k * w)] = count | ((uint32_t)(uint8_t)b8 << 24);
                // Occluder z-write for solid tiles (renderer/src/builder.rs:1014-1028; twin:
                // shaders/d3d11/propagate.cs.glsl:204-208). The fill rule is ignored, as in the reference.
                if (z_write && count == 0 && b8 != 0 && clip_ref == 0 && fx_ok) {
                    const int fy = path.min_y + y - b.fb.min_y;
                    if (fy >= 0 && fy < fb_h) atomicMax(z_buffer + (size_t)fy * fb_w + fx, (int32_t)path.global_path_id);
                }
                backdrop += delta; // backdrops[column] += delta (i32)   (renderer/src/tiler.rs:161)
            }
        }
        t += (uint32_t)(AHEAD * w);
    }
}

int launch_propagate(const BatchDev &b, uint32_t *tile_word, const int32_t *col_backdrop, int32_t *z_buffer,
                     const ClipDev *clip, uint32_t *tile_clip, uint32_t *tile_orig_count, cudaStream_t stream) {
    if (b.n_columns == 0) return 0;
    const bool dense = b.n_columns >= PROPAGATE_DENSE_COLUMNS;
    const unsigned grid = div_up(b.n_columns, 128);
    if (clip && tile_clip) {
        if (dense)
            launch_chained(k_propagate<true, true>, grid, 128, stream, b, tile_word, col_backdrop, z_buffer, *clip, tile_clip, tile_orig_count);
        else
            launch_chained(k_propagate<true, false>, grid, 128, stream, b, tile_word, col_backdrop, z_buffer, *clip, tile_clip, tile_orig_count);
    } else if (dense) {
        launch_chained(k_propagate<false, true>, grid, 128, stream, b, tile_word, col_backdrop, z_buffer, ClipDev{}, nullptr, nullptr);
    } else {
        launch_chained(k_propagate<false, false>, grid, 128, stream, b, tile_word, col_backdrop, z_buffer, ClipDev{}, nullptr, nullptr);
    }
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// ---------------------------------------------------------------------------------------------
// sort — per-framebuffer-tile painter's-order lists with z-cull.
//
// Count -> scan -> emit again: every surviving tile bumps its framebuffer tile's counter, the scan
// gives each framebuffer tile a contiguous run, and the emit pass appends 32-byte entries through
// a per-framebuffer-tile cursor. The append order inside a run is arbitrary; the fused fill+tile
// kernel sorts its (short) run by tile index — tiles are allocated path by path, so ascending
// tile index is draw order — in shared memory. This replaces the linked-list insertion sort of
// shaders/d3d11/sort.cs.glsl:60-95 and needs no global sort.
// ---------------------------------------------------------------------------------------------

// Also allocates the tile-grouped fill runs: a block sums the fill counts of its surviving tiles,
// reserves that many slots from a global cursor with ONE atomic, and hands each tile its run start
// from a block-local scan. Runs of different blocks land in arbitrary order, which nothing depends
// on (every reader goes through tile_fill_pos / TileEntry.fill_end) — and the device-wide scan over
// all bbox tiles that used to do this is gone.
constexpr int LIST_ITEMS = 8; // tiles per thread
constexpr int LIST_TILE = 256 * LIST_ITEMS;

#ifndef PF_LIST_MIN_BLOCKS
#define PF_LIST_MIN_BLOCKS 8 // 32 registers, full occupancy (random100k@8192 sort 0.120 -> 0.102 ms)
#endif
__global__ void __launch_bounds__(256, PF_LIST_MIN_BLOCKS)
    k_list_count(BatchDev b, const uint32_t *__restrict__ tile_word, const int32_t *__restrict__ z_buffer,
                 uint32_t *__restrict__ tile_fb, uint32_t *__restrict__ fb_count,
                 uint32_t *__restrict__ tile_fill_pos, uint32_t *__restrict__ fill_cursor,
                 uint32_t *__restrict__ path_live, int keep_all_fills, const uint32_t *__restrict__ run_counts,
                 uint32_t *__restrict__ live_tiles, uint32_t live_capacity, uint32_t *__restrict__ live_count,
                 uint32_t *__restrict__ fb_alpha, const uint32_t *__restrict__ tile_clip) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    __shared__ uint32_t smem[256 / 32 + 1];
    __shared__ uint32_t s_base;
    const uint32_t base = blockIdx.x * LIST_TILE + threadIdx.x;
    const int fb_w = b.fb.max_x - b.fb.min_x, fb_h = b.fb.max_y - b.fb.min_y;
    uint32_t run[LIST_ITEMS];
    uint32_t total = 0, live_mask = 0;
#pragma unroll
    for (int r = 0; r < LIST_ITEMS; r++) {
        const uint32_t t = base + (uint32_t)r * 256u;
        run[r] = 0;
        if (t >= b.n_tiles) continue;
        const uint32_t word = __ldg(tile_word + t);
        const uint32_t count = word & 0x00ffffffu;
        uint32_t result = 0xffffffffu;
        // Empty tiles are never drawn (renderer/src/builder.rs:1014-1016).
        if (count != 0 || (word >> 24) != 0) {
            const uint32_t p = search_coarse(b.path_tile_offset, b.tile_index, t);
            const PathInfo path = load_path(b.paths, p);
            const int w = path.max_x - path.min_x;
            const uint32_t local = t - path.tile_offset;
            // local = y * w + x without the ~60 instructions of two 32-bit divisions: a float quotient, corrected
            // (rects of more than 2^20 tiles take the integer path)
            uint32_t qy, qx;
            if (local < (1u << 20)) { // (quotient error of the approximate division < 0.25: at most one off before the correction)
                qy = (uint32_t)__float2uint_rz(__fdividef((float)local + 0.5f, (float)w));
                qx = local - qy * (uint32_t)w;
                if ((int32_t)qx < 0) qy--, qx += (uint32_t)w;
                else if (qx >= (uint32_t)w) qy++, qx -= (uint32_t)w;
            } else {
                qy = local / (uint32_t)w, qx = local - qy * (uint32_t)w;
            }
            const int x = (int)qx, y = (int)qy;
            const int fx = path.min_x + x - b.fb.min_x, fy = path.min_y + y - b.fb.min_y;
            if (fx >= 0 && fy >= 0 && fx < fb_w && fy < fb_h) {
                const uint32_t fbi = (uint32_t)(fy * fb_w + fx);
                // z-cull: dropped iff path_id < z (shaders/d3d11/sort.cs.glsl:74, d3d9/tile.vs.glsl:52-56)
                if ((int32_t)path.global_path_id >= __ldg(z_buffer + fbi)) {
                    result = fbi;
                    live_mask |= 1u << r;
                    atomicAdd(fb_count + fbi, 1u);
                    if (count != 0) path_live[p] = 1u; // some fills of this path will be read: its lines must be walked again
                    // The framebuffer tile needs per-pixel work (a mask of fills or a clip mask): plain store,
                    // every writer stores the same value.
                    if (count != 0 || (path.paint_ctrl & PATH_TEXTURED) || (tile_clip && __ldg(tile_clip + t) != 0u))
                        fb_alpha[fbi] = 1u;
                }
            }
        }
        tile_fb[t] = result;
        // Occlusion culling before fill emission: culled tiles get no run (parity dumps keep all).
        // (parity dumps with clipped paths: tiles dropped by the clip keep the run of their original fills)
        run[r] = (keep_all_fills || result != 0xffffffffu) ? (run_counts ? __ldg(run_counts + t) & 0x00ffffffu : count) : 0u;
        total += run[r];
    }
    uint32_t block_total;
    uint32_t pos = block_exclusive_scan(total, smem, &block_total);
    if (threadIdx.x == 0) s_base = block_total ? atomicAdd(fill_cursor, block_total) : 0u;
    __syncthreads();
    pos += s_base;
#pragma unroll
    for (int r = 0; r < LIST_ITEMS; r++) {
        if (run[r]) {
            tile_fill_pos[base + (uint32_t)r * 256u] = pos;
            pos += run[r];
        }
    }
    // Steady state: the surviving tiles are also written out as a compact list (arbitrary order), so the
    // entry pass visits those few instead of every bbox tile (random100k@8K: 0.76 M of 16.6 M).
    if (live_tiles) {
        // warp-level: a shuffle scan of the lanes' survivor counts and one counter increment per warp
        const int lane = threadIdx.x & 31;
        const uint32_t mine = (uint32_t)__popc(live_mask);
        uint32_t inclusive = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inclusive, d);
            if (lane >= d) inclusive += up;
        }
        const uint32_t warp_total = __shfl_sync(0xffffffffu, inclusive, 31);
        uint32_t warp_base = 0;
        if (lane == 31 && warp_total) warp_base = atomicAdd(live_count, warp_total);
        warp_base = __shfl_sync(0xffffffffu, warp_base, 31);
        uint32_t slot = warp_base + inclusive - mine;
#pragma unroll
        for (int r = 0; r < LIST_ITEMS; r++) {
            if (live_mask & (1u << r)) {
                if (slot < live_capacity) live_tiles[slot] = base + (uint32_t)r * 256u;
                slot++;
            }
        }
    }
}

int launch_list_count(const BatchDev &b, const uint32_t *tile_word, const int32_t *z_buffer, uint32_t *tile_fb,
                      uint32_t *fb_count, uint32_t *tile_fill_pos, uint32_t *fill_cursor, uint32_t *path_live,
                      bool keep_all_fills, const uint32_t *run_counts, uint32_t *live_tiles, uint32_t live_capacity,
                      uint32_t *live_count, uint32_t *fb_alpha, const uint32_t *tile_clip, cudaStream_t stream) {
    if (b.n_tiles == 0) return 0;
    launch_chained(k_list_count, div_up(b.n_tiles, LIST_TILE), 256, stream, b, tile_word, z_buffer, tile_fb, fb_count,
                   tile_fill_pos, fill_cursor, path_live, keep_all_fills ? 1 : 0, run_counts, live_tiles, live_capacity,
                   live_count, fb_alpha, tile_clip);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

template <bool COMPACT>
__global__ void __launch_bounds__(256)
    k_list_emit(BatchDev b, const uint32_t *__restrict__ tile_fb, const uint32_t *__restrict__ tile_word,
                const uint32_t *__restrict__ tile_fill_pos, const uint32_t *__restrict__ fb_start,
                uint32_t *__restrict__ fb_cursor, const float4 *__restrict__ paints,
                TileEntry *__restrict__ entries, uint32_t capacity, OverflowGuard guard, ClipDev clip,
                const uint32_t *__restrict__ tile_clip, uint2 *__restrict__ entry_clip,
                const uint32_t *__restrict__ live_tiles, uint32_t live_capacity, const uint32_t *__restrict__ live_count) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    // All totals are final by now. A batch that overflowed a stage buffer must leave the destination
    // untouched (the exact-sized re-run may have to load it): park the fused kernel's work counter
    // past the end so its warps find nothing to do.
    if (t == 0 && (guard.totals[0] > guard.line_bound || guard.totals[2] > guard.entry_bound ||
                   guard.totals[5] > guard.fill_bound))
        *guard.work_counter = 0xf0000000u;
    bool in_range = t < b.n_tiles;
    if (COMPACT) { // thread i takes the i-th surviving tile
        in_range = t < min(__ldg(live_count), live_capacity);
        t = in_range ? __ldg(live_tiles + t) : 0u;
    }
    uint32_t fbi = in_range ? __ldg(tile_fb + t) : 0xffffffffu;
    const bool live = fbi != 0xffffffffu;
    if (!__any_sync(0xffffffffu, live)) return;
    uint32_t p = live ? search_coarse(b.path_tile_offset, b.tile_index, t) : 0;
    if (live) {
        TileEntry e;
        e.fill_end = __ldg(tile_fill_pos + t); // the bin emit pass left the cursor at the end of the run
        e.word = __ldg(tile_word + t);
        const uint32_t path_paint_ctrl = __ldg(&b.paths[p].paint_ctrl);
        e.paint_ctrl = (path_paint_ctrl & 0x00ffffffu) | ((path_paint_ctrl & PATH_TEXTURED) ? ENTRY_TEXTURED : 0u);
        e.tile_index = t;
        uint32_t slot = __ldg(fb_start + fbi) + atomicAdd(fb_cursor + fbi, 1u);
        // The paint premultiplied, (rgb * a, a): what the compositing kernels blend with (color.rgb *= color.a,
        // shaders/tile_fragment.inc.glsl:611).
        float4 color = __ldg(paints + (e.paint_ctrl & 0xffffu));
        color.x *= color.w, color.y *= color.w, color.z *= color.w;
        uint2 clip_entry = make_uint2(0, 0);
        if (tile_clip) { // the batch has clipped paths
            const uint32_t ref = __ldg(tile_clip + t);
            if (ref != 0) {
                const uint32_t ct = (ref & ~TILE_CLIP_REPLACE) - 1u;
                clip_entry = make_uint2(__ldg(clip.tile_fill_end + ct), __ldg(clip.tile_word + ct));
                e.paint_ctrl |= ENTRY_HAS_CLIP | ((ref & TILE_CLIP_REPLACE) ? ENTRY_CLIP_REPLACE : 0u);
            }
        }
        if (slot < capacity) {
            *reinterpret_cast<uint4 *>(entries + slot) = *reinterpret_cast<uint4 *>(&e);
            entries[slot].color = color;
            if (tile_clip) entry_clip[slot] = clip_entry;
        }
    }
}

int launch_list_emit(const BatchDev &b, const uint32_t *tile_fb, const uint32_t *tile_word,
                     const uint32_t *tile_fill_pos, const uint32_t *fb_start, uint32_t *fb_cursor,
                     const float4 *paints, TileEntry *entries, uint32_t capacity, const OverflowGuard &guard,
                     const ClipDev *clip, const uint32_t *tile_clip, uint2 *entry_clip, const uint32_t *live_tiles,
                     uint32_t live_capacity, const uint32_t *live_count, cudaStream_t stream) {
    if (b.n_tiles == 0) return 0;
    const ClipDev clip_dev = clip && tile_clip ? *clip : ClipDev{};
    if (live_tiles)
        launch_chained(k_list_emit<true>, std::max(1u, div_up(live_capacity, 256)), 256, stream,
                       b, tile_fb, tile_word, tile_fill_pos, fb_start, fb_cursor, paints, entries, capacity, guard, clip_dev,
                       clip ? tile_clip : nullptr, entry_clip, live_tiles, live_capacity, live_count);
    else
        launch_chained(k_list_emit<false>, div_up(b.n_tiles, 256), 256, stream,
                       b, tile_fb, tile_word, tile_fill_pos, fb_start, fb_cursor, paints, entries, capacity, guard, clip_dev,
                       clip ? tile_clip : nullptr, entry_clip, (const uint32_t *)nullptr, 0u, (const uint32_t *)nullptr);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// ---------------------------------------------------------------------------------------------
// fill + tile (fused) live in composite.cu. What follows is the coverage evaluation kept for the
// parity dumps (k_alpha_masks): computeCoverage in the form round 1 shipped, one lane per
// (column, 8-row half) — an independent evaluation of the same reference math.
// ---------------------------------------------------------------------------------------------

// Coverage contributions are accumulated as integers so the sum does not depend on the order of
// a tile's fills (the tile-grouped fill array is filled through an atomic cursor): adding
// 1.5 * 2^8 pins the float's exponent so its mantissa is the contribution in units of 2^-15.
constexpr float COV_MAGIC = 384.0f;
constexpr uint32_t COV_MAGIC_BITS = 0x43c00000u;
constexpr float COV_SCALE = 1.0f / 32768.0f;

// computeCoverage (shaders/fill_area.inc.glsl:11-27) for one pixel column and NS vertically
// adjacent 4-row strips: (cx, cy) is the centre of the first pixel of the first strip in tile
// space. Everything that depends only on x (window, dX, t) is computed once per column; each strip
// costs one LUT fetch (4 rows per texel) and four multiply-adds. Returns false when the column is
// outside the segment's x range (the LUT term is multiplied by dX = 0 in the reference).
template <int NS>
__device__ __forceinline__ bool accumulate_fill(uint32_t from_w, uint32_t to_w, float cx, float cy,
                                                cudaTextureObject_t lut, uint32_t (&acc)[NS][4]) {
    const float s = 1.0f / 256.0f;
    float from_x = fmaf((float)(from_w & 0xffffu), s, -cx), to_x = fmaf((float)(to_w & 0xffffu), s, -cx);
    float wx = fminf(fmaxf(from_x, -0.5f), 0.5f), wy = fminf(fmaxf(to_x, -0.5f), 0.5f);
    float dX = wx - wy;
    if (dX == 0.0f) return false;
    float from_y = fmaf((float)(from_w >> 16), s, -cy), to_y = fmaf((float)(to_w >> 16), s, -cy);
    bool from_left = from_x < to_x;
    float lx = from_left ? from_x : to_x, ly = from_left ? from_y : to_y;
    float rx = from_left ? to_x : from_x, ry = from_left ? to_y : from_y;
    float inv = __fdividef(1.0f, rx - lx);
    float t = (0.5f * (wx + wy) - lx) * inv;
    float y = fmaf(ry - ly, t, ly);
    float v = fabsf((ry - ly) * inv * dX) * (1.0f / 16.0f);
#pragma unroll
    for (int k = 0; k < NS; k++) {
        // the strip k rows lower sees the segment 4k px higher: y - 4k
        float4 tex = tex2D<float4>(lut, (y + (8.0f - 4.0f * (float)k)) * (1.0f / 16.0f), v);
        acc[k][0] += __float_as_uint(fmaf(tex.x, dX, COV_MAGIC));
        acc[k][1] += __float_as_uint(fmaf(tex.y, dX, COV_MAGIC));
        acc[k][2] += __float_as_uint(fmaf(tex.z, dX, COV_MAGIC));
        acc[k][3] += __float_as_uint(fmaf(tex.w, dX, COV_MAGIC));
    }
    return true;
}

__device__ __forceinline__ float finish_coverage(uint32_t acc, uint32_t contributions) {
    return (float)(int32_t)(acc - contributions * COV_MAGIC_BITS) * COV_SCALE;
}

// ---------------------------------------------------------------------------------------------
// Parity-dump helpers: alpha tile numbering in SequentialExecutor order, D3D9-form lists.
// ---------------------------------------------------------------------------------------------

// A tile's alpha id is allocated by its first surviving fill (renderer/src/builder.rs:555-576), so
// ids in first-fill order = exclusive scan, over emission order, of "this fill is its tile's first".
__global__ void k_alpha_flags(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_first_fill,
                              uint8_t *fill_is_first) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    if ((tile_word[t] & 0x00ffffffu) != 0) fill_is_first[tile_first_fill[t]] = 1;
}
int launch_alpha_flags(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_first_fill,
                       uint8_t *fill_is_first, cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    k_alpha_flags<<<div_up(n_tiles, 256), 256, 0, stream>>>(n_tiles, tile_word, tile_first_fill, fill_is_first);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

__global__ void k_alpha_assign(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_first_fill,
                               const uint32_t *fill_first_scan, uint32_t *tile_alpha_id, uint32_t alpha_base) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    tile_alpha_id[t] = (tile_word[t] & 0x00ffffffu) != 0 ? alpha_base + fill_first_scan[tile_first_fill[t]] : 0xffffffffu;
}
int launch_alpha_assign(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_first_fill,
                        const uint32_t *fill_first_scan, uint32_t *tile_alpha_id, uint32_t alpha_base,
                        cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    k_alpha_assign<<<div_up(n_tiles, 256), 256, 0, stream>>>(n_tiles, tile_word, tile_first_fill, fill_first_scan,
                                                             tile_alpha_id, alpha_base);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

struct FillRecord { // gpu_data.rs:354-363
    uint16_t from_x, from_y, to_x, to_y;
    uint32_t link;
};
__global__ void k_dump_fills(uint32_t n, const EmitFill *fills_emit, const uint32_t *tile_alpha_id, FillRecord *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EmitFill f = fills_emit[i];
    out[i] = FillRecord{(uint16_t)(f.from & 0xffff), (uint16_t)(f.from >> 16), (uint16_t)(f.to & 0xffff),
                        (uint16_t)(f.to >> 16), tile_alpha_id[f.tile]};
}
int launch_dump_fills(uint32_t n_fills, const EmitFill *fills_emit, const uint32_t *tile_alpha_id, void *out,
                      cudaStream_t stream) {
    if (n_fills == 0) return 0;
    k_dump_fills<<<div_up(n_fills, 256), 256, 0, stream>>>(n_fills, fills_emit, tile_alpha_id, (FillRecord *)out);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

__global__ void k_dump_tile_flags(uint32_t n_tiles, const uint32_t *tile_word, uint32_t *flags) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    uint32_t w = tile_word[t];
    flags[t] = ((w & 0x00ffffffu) != 0 || (w >> 24) != 0) ? 1u : 0u;
}
int launch_dump_tile_flags(uint32_t n_tiles, const uint32_t *tile_word, uint32_t *flags, cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    k_dump_tile_flags<<<div_up(n_tiles, 256), 256, 0, stream>>>(n_tiles, tile_word, flags);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

struct TileRecord { // TileObjectPrimitive, gpu_data.rs:264-275
    int16_t tile_x, tile_y;
    uint32_t alpha_tile_id;
    uint32_t path_id;
    uint16_t color;
    uint8_t ctrl;
    int8_t backdrop;
};
__global__ void k_dump_tiles(BatchDev b, const uint32_t *tile_word, const uint32_t *tile_alpha_id,
                             const uint32_t *flags, const uint32_t *pos, TileRecord *out, const uint32_t *tile_clip,
                             const uint32_t *clip_tile_word, const uint32_t *clip_alpha_id) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.n_tiles || !flags[t]) return;
    uint32_t p = search_coarse(b.path_tile_offset, b.tile_index, t);
    const PathInfo path = load_path(b.paths, p);
    int w = path.max_x - path.min_x;
    uint32_t local = t - path.tile_offset;
    TileRecord r;
    r.tile_x = (int16_t)(path.min_x + (int)(local % (uint32_t)w));
    r.tile_y = (int16_t)(path.min_y + (int)(local / (uint32_t)w));
    r.alpha_tile_id = tile_alpha_id[t];
    r.path_id = path.global_path_id;
    r.color = (uint16_t)(path.paint_ctrl & 0xffff);
    r.ctrl = (uint8_t)((path.paint_ctrl >> 16) & 0xff);
    r.backdrop = (int8_t)(tile_word[t] >> 24);
    if (tile_clip && tile_clip[t] != 0) { // Tiler::prepare_tiles, renderer/src/tiler.rs:124-141
        const uint32_t ref = tile_clip[t], ct = (ref & ~TILE_CLIP_REPLACE) - 1u;
        if (ref & TILE_CLIP_REPLACE) {
            r.alpha_tile_id = clip_alpha_id[ct];
            r.backdrop = (int8_t)(clip_tile_word[ct] >> 24);
        } else {
            r.backdrop = 0; // moved into the Clip record's dest_backdrop
        }
    }
    out[pos[t]] = r;
}

// Clip records of the D3D9 batch (gpu_data.rs Clip; builder.rs:1031-1040): one per draw tile whose mask is
// combined with a clip tile's mask, in tile order.
struct ClipRecord {
    uint32_t dest_tile_id;
    int32_t dest_backdrop;
    uint32_t src_tile_id;
    int32_t src_backdrop;
};
__global__ void k_dump_clip_flags(uint32_t n_tiles, const uint32_t *tile_clip, uint32_t *flags) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    flags[t] = (tile_clip[t] != 0 && !(tile_clip[t] & TILE_CLIP_REPLACE)) ? 1u : 0u;
}
__global__ void k_dump_clips(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_alpha_id,
                             const uint32_t *tile_clip, const uint32_t *clip_tile_word, const uint32_t *clip_alpha_id,
                             const uint32_t *flags, const uint32_t *pos, ClipRecord *out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles || !flags[t]) return;
    const uint32_t ct = tile_clip[t] - 1u;
    out[pos[t]] = ClipRecord{tile_alpha_id[t], (int32_t)(int8_t)(tile_word[t] >> 24), clip_alpha_id[ct],
                             (int32_t)(int8_t)(clip_tile_word[ct] >> 24)};
}
int launch_dump_clip_flags(uint32_t n_tiles, const uint32_t *tile_clip, uint32_t *flags, cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    k_dump_clip_flags<<<div_up(n_tiles, 256), 256, 0, stream>>>(n_tiles, tile_clip, flags);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}
int launch_dump_clips(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_alpha_id, const uint32_t *tile_clip,
                      const uint32_t *clip_tile_word, const uint32_t *clip_alpha_id, const uint32_t *flags,
                      const uint32_t *pos, void *out, cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    k_dump_clips<<<div_up(n_tiles, 256), 256, 0, stream>>>(n_tiles, tile_word, tile_alpha_id, tile_clip, clip_tile_word,
                                                           clip_alpha_id, flags, pos, (ClipRecord *)out);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}
int launch_dump_tiles(const BatchDev &b, const uint32_t *tile_word, const uint32_t *tile_alpha_id,
                      const uint32_t *flags, const uint32_t *pos, void *out, const uint32_t *tile_clip,
                      const uint32_t *clip_tile_word, const uint32_t *clip_alpha_id, cudaStream_t stream) {
    if (b.n_tiles == 0) return 0;
    k_dump_tiles<<<div_up(b.n_tiles, 256), 256, 0, stream>>>(b, tile_word, tile_alpha_id, flags, pos, (TileRecord *)out,
                                                             tile_clip, clip_tile_word, clip_alpha_id);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// Coverage masks per alpha tile with the same device function the fused kernel uses.
__global__ void __launch_bounds__(32)
    k_alpha_masks(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_fill_pos,
                  const uint32_t *tile_alpha_id, const PackedFill *fills, cudaTextureObject_t lut, float *out) {
    const int x = threadIdx.x & 15, half = threadIdx.x >> 4;
    const float cx = (float)x + 0.5f, cy = (float)(half * 8) + 0.5f;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        uint32_t count = tile_word[t] & 0x00ffffffu;
        if (count == 0) continue;
        uint32_t end = tile_fill_pos[t];
        uint32_t acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        uint32_t contributions = 0;
        for (uint32_t fi = end - count; fi < end; fi++) {
            PackedFill f = fills[fi];
            contributions += accumulate_fill<2>(f.x, f.y, cx, cy, lut, acc) ? 1u : 0u;
        }
        float *mask = out + (size_t)tile_alpha_id[t] * 256;
#pragma unroll
        for (int k = 0; k < 8; k++) mask[(half * 8 + k) * 16 + x] = finish_coverage(acc[k >> 2][k & 3], contributions);
    }
}
int launch_alpha_masks(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_fill_pos,
                       const uint32_t *tile_alpha_id, const PackedFill *fills, cudaTextureObject_t area_lut,
                       float *out, cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    unsigned grid = n_tiles < 16384u ? n_tiles : 16384u;
    k_alpha_masks<<<grid, 32, 0, stream>>>(n_tiles, tile_word, tile_fill_pos, tile_alpha_id, fills, area_lut, out);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

} // namespace pf
