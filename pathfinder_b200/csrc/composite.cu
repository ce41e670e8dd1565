// pathfinder_b200/csrc/composite.cu — fill + tile (fused): exact-area coverage from the LUT, fill rule,
// paint, SrcOver blend, RGBA8 store. The alpha mask never exists in HBM.
//
// Reference math restated here (paths relative to the reference checkout):
//   fill   shaders/fill_area.inc.glsl:11-27 (computeCoverage), shaders/d3d11/fill_compute.inc.glsl:11-25
//   tile   shaders/d3d11/tile.cs.glsl:91-159 (painter's-order loop, dest * (1 - a) + src),
//          shaders/tile_fragment.inc.glsl:539-614 (sampleMask, calculateColor),
//          shaders/d3d9/tile_clip_combine.fs.glsl:28-31 (clip mask combine)
//
// Two kernels instead of one warp per framebuffer tile for everything (round 1):
//   k_tile_solid   one LANE per framebuffer tile. A tile whose list holds only solid entries (no fills,
//                  no clip mask) has a single colour: the lane blends the few entries in draw order and the
//                  warp then writes its 32 tiles row by row with 512-byte coalesced 128-bit stores. Tiles
//                  that need per-pixel work are appended to a queue.
//   k_tile_alpha   persistent warps pull queued tiles. Between two tiles with fills, solid entries are
//                  folded into one affine map (d -> d * s + o), so only entries with fills touch the 256
//                  pixels. Per fill, everything that does not depend on the pixel column is computed once
//                  by the lane that loaded the fill and broadcast through shared memory; the pixel state
//                  lives in shared memory while the fills are evaluated, so the loop runs at low register
//                  count. Finished tiles are transposed through shared memory and leave as 128-bit stores.
// This file is compiled with FMA contraction on (nothing here is compared bit-exactly with the CPU tiler;
// the bar is 1/255 per channel), unlike kernels.cu.
#include "common.cuh"
#include "kernels.cuh"

namespace pf {

namespace {

#ifndef PF_TILE_WARPS
#define PF_TILE_WARPS 4 // warps per block of k_tile_alpha
#endif
#ifndef PF_TILE_MIN_BLOCKS
#define PF_TILE_MIN_BLOCKS 7 // resident blocks per SM the register allocation must allow (70 registers; 6 and 8 measured slower)
#endif
#ifndef PF_PREFETCH
#define PF_PREFETCH 1
#endif
#ifndef PF_FILL_FFMA2
#define PF_FILL_FFMA2 1
#endif
#ifndef PF_SIMPLE_STORE
#define PF_SIMPLE_STORE 1
#endif

constexpr int TILE_WARPS = PF_TILE_WARPS;
#ifndef PF_ENTRY_CAP
#define PF_ENTRY_CAP 32
#endif
constexpr int ENTRY_CAP = PF_ENTRY_CAP;  // entries rank-sorted in shared memory; deeper lists are walked by selection
constexpr int SOLID_MAX = 8;   // deepest all-solid list k_tile_solid blends itself
constexpr uint32_t PF_FILTER_TEXT_KIND = 2; // PF_FILTER_TEXT (include/pf_cuda.h)
constexpr uint32_t WORK_PARKED = 0xf0000000u; // k_list_emit parks the tile counter here when a stage overflowed

// Coverage contributions are accumulated as integers so the sum does not depend on the order of a tile's
// fills (the tile-grouped fill array is filled through an atomic cursor): adding 1.5 * 2^8 pins the
// float's exponent so its mantissa is the contribution in units of 2^-15.
constexpr float COV_MAGIC = 384.0f;
constexpr uint32_t COV_MAGIC_BITS = 0x43c00000u;
constexpr float COV_SCALE = 1.0f / 32768.0f;

// sampleMask (shaders/tile_fragment.inc.glsl:539-556) on coverage = mask + backdrop.
__device__ __forceinline__ float mask_alpha(float coverage, uint32_t ctrl) {
    if (ctrl & 1u) { // TILE_CTRL_MASK_WINDING
        coverage = fabsf(coverage);
    } else if (ctrl & 2u) { // TILE_CTRL_MASK_EVEN_ODD
        float m = coverage - 2.0f * floorf(coverage * 0.5f);
        coverage = 1.0f - fabsf(1.0f - m);
    } else {
        coverage = 1.0f;
    }
    return fminf(1.0f, coverage);
}

// Packed f32x2 arithmetic (sm_100: FFMA2 / FMUL2 issue two fp32 operations per lane per instruction).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// A pixel as two packed pairs (r, g), (b, a).
struct Px {
    f32x2 rg, ba;
};
__device__ __forceinline__ Px px_from(float4 c) { return Px{pack2(c.x, c.y), pack2(c.z, c.w)}; }
__device__ __forceinline__ float4 px_to(Px p) {
    float4 c;
    unpack2(p.rg, c.x, c.y);
    unpack2(p.ba, c.z, c.w);
    return c;
}

// dest = dest * (1 - a) + src (shaders/d3d11/tile.cs.glsl:155) with src = (rgb * a, a), a = paint alpha * mask:
// with the paint premultiplied, P = (rgb * w, w), this is d + m * (P - w * d) — two packed FMAs per pair.
__device__ __forceinline__ void over(Px &d, Px paint, f32x2 neg_w, float m) {
    const f32x2 mm = pack2(m, m);
    d.rg = fma2(mm, fma2(neg_w, d.rg, paint.rg), d.rg);
    d.ba = fma2(mm, fma2(neg_w, d.ba, paint.ba), d.ba);
}

// clamp-free x * 255 rounded to nearest even in the low mantissa bits of x * 255 + 2^23 (no F2I on the slow
// pipe); colours are convex combinations of values in [0, 1], so only rounding noise can leave the range
// and it vanishes in the rounding. The four bytes are packed with PRMT.
__device__ __forceinline__ uint32_t pack_rgba8(Px p) {
    const f32x2 scale = pack2(255.0f, 255.0f), magic = pack2(8388608.0f, 8388608.0f);
    float r, g, b, a;
    unpack2(fma2(p.rg, scale, magic), r, g);
    unpack2(fma2(p.ba, scale, magic), b, a);
    const uint32_t rg = __byte_perm(__float_as_uint(r), __float_as_uint(g), 0x0040); // bytes: r0, g0, 0, 0
    const uint32_t ba = __byte_perm(__float_as_uint(b), __float_as_uint(a), 0x0040);
    return __byte_perm(rg, ba, 0x5410);
}

__device__ __forceinline__ float4 unpack_rgba8(uint32_t v) {
    const float s = 1.0f / 255.0f;
    return make_float4((float)(v & 0xff) * s, (float)((v >> 8) & 0xff) * s, (float)((v >> 16) & 0xff) * s,
                       (float)(v >> 24) * s);
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Column-independent part of computeCoverage for one fill, computed once by the lane that loaded it.
struct __align__(16) FillParams {
    float lx, rx;   // x of the left / right end point, tile space [0, 16]
    float ly05;     // y of the left end point minus 0.5 (the centre of pixel row 0)
    float dy;       // right.y - left.y
    float inv;      // 1 / (right.x - left.x)
    float slope16;  // |dy / dx| / 16: the LUT's second coordinate per unit of window width
    float sign;     // sign of dX = window(from).x - window(to).x: -1 when `from` is the left end
    float pad;
};

__device__ __forceinline__ FillParams fill_params(uint2 fill) {
    const float s = 1.0f / 256.0f;
    const float fx = (float)(fill.x & 0xffffu) * s, fy = (float)(fill.x >> 16) * s;
    const float tx = (float)(fill.y & 0xffffu) * s, ty = (float)(fill.y >> 16) * s;
    const bool from_left = fx < tx;
    FillParams p;
    p.lx = from_left ? fx : tx;
    p.rx = from_left ? tx : fx;
    const float ly = from_left ? fy : ty, ry = from_left ? ty : fy;
    p.ly05 = ly - 0.5f;
    p.dy = ry - ly;
    p.inv = __fdividef(1.0f, p.rx - p.lx); // from_x != to_x (degenerate fills are culled by add_fill)
    p.slope16 = fabsf(p.dy * p.inv) * (1.0f / 16.0f);
    p.sign = from_left ? -1.0f : 1.0f;
    p.pad = 0.0f;
    return p;
}

// ---------------------------------------------------------------------------------------------
// k_tile_solid — single-colour tiles, one lane per framebuffer tile; queues the others.
// ---------------------------------------------------------------------------------------------

#ifndef PF_SOLID_MIN_BLOCKS
#define PF_SOLID_MIN_BLOCKS 1
#endif
template <bool LOAD_DEST>
__global__ void __launch_bounds__(128, PF_SOLID_MIN_BLOCKS) k_tile_solid(CompositeArgs a) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int fb_w = a.fb.max_x - a.fb.min_x;
    const uint32_t segs = ((uint32_t)fb_w + 31u) >> 5; // 32-tile row segments per tile row
    const uint32_t rows = (uint32_t)(a.tile_y1 - a.tile_y0);
    if (warp_global >= segs * rows) return;
    if (*a.work_counter >= WORK_PARKED) return; // a stage overflowed: the batch is re-run, leave the image alone
    const uint32_t tile_row = warp_global / segs, seg = warp_global - tile_row * segs;
    const int col = (int)(seg * 32u) + lane;
    const int ty = a.tile_y0 + (int)tile_row;
    const bool in_row = col < fb_w;
    uint32_t n = 0, e0 = 0;
    bool has_alpha = false;
    if (in_row && ty >= a.fb.min_y && ty < a.fb.max_y) {
        const size_t index = (size_t)(ty - a.fb.min_y) * (size_t)fb_w + (size_t)col;
        // (three independent loads: no load waits for another)
        n = __ldg(a.fb_count + index);
        e0 = __ldg(a.fb_start + index);
        has_alpha = __ldg(a.fb_alpha + index) != 0u;
    }
    // Tiles that need per-pixel work go to the queue of k_tile_alpha (one counter increment per warp).
    const bool queued = in_row && (has_alpha || n > (uint32_t)SOLID_MAX || (LOAD_DEST && n > 0));
    const uint32_t queued_mask = __ballot_sync(0xffffffffu, queued);
    if (queued_mask) {
        uint32_t base = 0;
        const int leader = __ffs(queued_mask) - 1;
        if (lane == leader) base = atomicAdd(a.queue_count, (uint32_t)__popc(queued_mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (queued) {
            const uint32_t slot = base + (uint32_t)__popc(queued_mask & ((1u << lane) - 1u));
            a.queue[slot] = (tile_row << 16) | (uint32_t)col;
            a.queue_hdr[slot] = make_uint2(n, e0); // (saves k_tile_alpha a dependent load per tile)
        }
    }
    // LOAD_ACTION_LOAD: a tile without entries keeps what the previous batches drew.
    const bool paints = in_row && !queued && !(LOAD_DEST && n == 0);
    const uint32_t paint_mask = __ballot_sync(0xffffffffu, paints);
    if (paint_mask == 0) {
        if (a.export_solid_mask && lane == 0) a.export_solid_mask[warp_global] = 0u;
        return;
    }

    // The tile's colour: its (few, solid) entries blended in draw order = ascending tile index, selected by
    // repeated minimum (the run is in arbitrary order).
    Px c = px_from(a.clear_color);
    if (paints) {
        uint32_t keys[SOLID_MAX]; // one round of loads for all the keys, then selection in registers
#pragma unroll
        for (int j = 0; j < SOLID_MAX; j++) // (a list of one needs no order: no round trip for its key)
            keys[j] = ((uint32_t)j < n && n > 1u) ? __ldg(&a.entries[e0 + j].tile_index) : ((uint32_t)j < n ? 0u : 0xffffffffu);
        uint32_t last = 0;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t best = 0xffffffffu, best_j = 0;
#pragma unroll
            for (int j = 0; j < SOLID_MAX; j++)
                if ((i == 0 || keys[j] > last) && keys[j] < best) best = keys[j], best_j = (uint32_t)j;
            last = best;
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.entries + e0 + best_j));
            const float4 paint = __ldg(&a.entries[e0 + best_j].color);
            // Solid tile: coverage = backdrop for every pixel (tile_fragment.inc.glsl:548).
            const float m = mask_alpha((float)(int)(int8_t)(raw.y >> 24), (raw.z >> 16) & 0xffu);
            over(c, px_from(paint), pack2(-paint.w, -paint.w), m);
        }
    }
    const uint32_t packed = pack_rgba8(c);
    if (a.export_solid_color) { // the other ranks expand these tiles themselves from 4 bytes each
        a.export_solid_color[(size_t)warp_global * 32u + (uint32_t)lane] = packed;
        if (lane == 0) a.export_solid_mask[warp_global] = paint_mask;
    }

    // Store: image row r of the segment's 32 tiles is 2 KB of consecutive bytes; lane l writes the 16-byte
    // chunks l, l + 32, l + 64, l + 96 of it, chunk c belonging to tile c / 4.
    uint32_t chunk_color[4];
    bool chunk_on[4], chunk_full[4];
    int chunk_px[4];
    const int seg_tx = a.fb.min_x + (int)(seg * 32u);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int j = q * 8 + (lane >> 2);
        chunk_color[q] = __shfl_sync(0xffffffffu, packed, j);
        chunk_px[q] = (seg_tx + j) * 16 + (lane & 3) * 4;
        chunk_on[q] = ((paint_mask >> j) & 1u) && chunk_px[q] + 4 > 0 && chunk_px[q] < a.dest_w;
        chunk_full[q] = chunk_px[q] >= 0 && chunk_px[q] + 4 <= a.dest_w && (a.dest_align_mask & 15u) == 0;
    }
    const int y_lo = max(0, ty * 16), y_hi = min(a.dest_h, ty * 16 + 16);
    for (int d = 0; d < a.n_dest; d++) {
        uint8_t *row = a.dests[d] + (ptrdiff_t)y_lo * (ptrdiff_t)a.dest_pitch;
        for (int y = y_lo; y < y_hi; y++, row += a.dest_pitch) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (!chunk_on[q]) continue;
                if (chunk_full[q]) {
                    *reinterpret_cast<uint4 *>(row + (ptrdiff_t)chunk_px[q] * 4) =
                        make_uint4(chunk_color[q], chunk_color[q], chunk_color[q], chunk_color[q]);
                } else {
                    for (int k = 0; k < 4; k++) {
                        const int px = chunk_px[q] + k;
                        if (px >= 0 && px < a.dest_w) *reinterpret_cast<uint32_t *>(row + (ptrdiff_t)px * 4) = chunk_color[q];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_tile_alpha — tiles with fills (or clip masks, or deep lists), one warp per tile.
// ---------------------------------------------------------------------------------------------

// computeCoverage (shaders/fill_area.inc.glsl:11-27) of one fill for pixel column `xf` (its left edge) and
// the two vertically adjacent 4-row strips this lane owns (u_off: LUT coordinate offset of the first).
// Branch-free: a column outside the fill's x range has dX = 0 and adds exactly COV_MAGIC, like the
// reference's `texture(...) * dX`, so every fill counts as one contribution on every lane.
// (Measured and dropped: summing in floats with a pinned exponent and round-down FMAs — exact and order-independent
// too, four FFMA2.RM or eight FFMA.RM per fill and no integer adds — is slower on sm_100a: 0.237 against 0.222 ms.)
struct CovAcc {
    uint32_t v[8];
};
__device__ __forceinline__ void acc_reset(CovAcc &acc) {
#pragma unroll
    for (int k = 0; k < 8; k++) acc.v[k] = 0u;
}
__device__ __forceinline__ void accumulate_fill(const float4 p0, const float4 p1, float xf, float u_off,
                                                cudaTextureObject_t lut, CovAcc &acc) {
    // p0 = {lx, rx, ly05, dy}, p1 = {inv, slope16, sign, -}
    const float l0 = p0.x - xf;                             // left.x relative to the column's left edge
    const float wl = __saturatef(l0), wr = __saturatef(p0.y - xf); // window = clamp(x, -0.5, 0.5) + 0.5
    const float width = wr - wl;
    const float dX = p1.z * width;
    const float t = fmaf(wl + wr, 0.5f, -l0) * p1.x;        // (mid(window) - left.x) / (right.x - left.x)
    const float y = fmaf(p0.w, t, p0.z);                    // mix(left.y, right.y, t), relative to row 0's centre
    const float v = p1.y * width;                           // abs(d * dX) / 16
    const float u = fmaf(y, 1.0f / 16.0f, u_off);           // (y + 8) / 16 for the first strip
    const float4 a0 = tex2D<float4>(lut, u, v);
    const float4 a1 = tex2D<float4>(lut, u - 0.25f, v);     // the strip 4 rows lower sees the segment 4 px higher
    // two packed FMAs per strip (the texture unit returns the four rows in consecutive registers): adding 1.5 * 2^8
    // pins the product's exponent, so its mantissa is the contribution in units of 2^-15 — summed as integers
    const f32x2 dd = pack2(dX, dX), magic = pack2(COV_MAGIC, COV_MAGIC);
    float c0, c1, c2, c3, c4, c5, c6, c7;
    unpack2(fma2(pack2(a0.x, a0.y), dd, magic), c0, c1);
    unpack2(fma2(pack2(a0.z, a0.w), dd, magic), c2, c3);
    unpack2(fma2(pack2(a1.x, a1.y), dd, magic), c4, c5);
    unpack2(fma2(pack2(a1.z, a1.w), dd, magic), c6, c7);
    acc.v[0] += __float_as_uint(c0);
    acc.v[1] += __float_as_uint(c1);
    acc.v[2] += __float_as_uint(c2);
    acc.v[3] += __float_as_uint(c3);
    acc.v[4] += __float_as_uint(c4);
    acc.v[5] += __float_as_uint(c5);
    acc.v[6] += __float_as_uint(c6);
    acc.v[7] += __float_as_uint(c7);
}

template <bool GENERAL>
struct __align__(16) TileWarpShared {
    float4 dst[8 * 32];         // the tile's pixels between two entries with fills: [row k of the lane][lane]
    uint4 entry[ENTRY_CAP];     // the run in draw order: {fill_end, count | backdrop, paint | ctrl | flags, tile index}
    float4 paint[ENTRY_CAP];    // premultiplied paint colour
    union {
        float4 fill[32][2];     // FillParams of up to 32 fills
        uint32_t stage[16 * 16 + 16]; // finished RGBA8 tile, row-major (+16 words for rows 8..15: conflict-free)
    };
    uint2 clip[GENERAL ? ENTRY_CAP : 1]; // {clip fill end, clip tile word} (batches with clipped paths)
};

// Coverage of the fills [end - count, end) plus the backdrop -> cov.
template <bool GENERAL>
__device__ __forceinline__ void fill_coverage(TileWarpShared<GENERAL> &sh, const PackedFill *__restrict__ fills,
                                              uint32_t end, uint32_t count,
                                              float backdrop, float xf, float u_off, cudaTextureObject_t lut, int lane,
                                              float (&cov)[8]) {
    CovAcc acc;
    acc_reset(acc);
    for (uint32_t f0 = end - count; f0 < end; f0 += 32) {
        const uint32_t m = min(32u, end - f0);
        if ((uint32_t)lane < m) {
            const FillParams p = fill_params(__ldg(fills + f0 + lane));
            sh.fill[lane][0] = make_float4(p.lx, p.rx, p.ly05, p.dy);
            sh.fill[lane][1] = make_float4(p.inv, p.slope16, p.sign, 0.0f);
        }
        __syncwarp();
#pragma unroll 2
        for (uint32_t j = 0; j < m; j++) accumulate_fill(sh.fill[j][0], sh.fill[j][1], xf, u_off, lut, acc);
        __syncwarp(); // the next batch (or the store staging) overwrites the parameters
    }
    // acc -> coverage for `count` contributions. Up to 127 contributions of magnitude <= 2^15 units: |sum| <
    // 2^22, so the signed sum can be read off the mantissa of 1.5 * 2^23 + sum (no I2F on the quarter-rate
    // pipe) and scaled, un-biased (12582912 * 2^-15 = 384) and offset by the backdrop in one FFMA.
    if (count <= 127u) {
        const uint32_t bias = 0x4b400000u - count * COV_MAGIC_BITS;
        const float offset = backdrop - 384.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) cov[k] = fmaf(__uint_as_float(acc.v[k] + bias), COV_SCALE, offset);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) cov[k] = (float)(int32_t)(acc.v[k] - count * COV_MAGIC_BITS) * COV_SCALE + backdrop;
    }
}

// coverage -> mask alpha by the tile's fill rule (sampleMask), eight pixels.
__device__ __forceinline__ void rule_of(float (&cov)[8], uint32_t ctrl) {
    if (ctrl & 1u) { // TILE_CTRL_MASK_WINDING
#pragma unroll
        for (int k = 0; k < 8; k++) cov[k] = fminf(fabsf(cov[k]), 1.0f);
    } else if (ctrl & 2u) { // TILE_CTRL_MASK_EVEN_ODD
        // 1 - |1 - (c mod 2)| is the distance from c to the nearest even integer: 2 * |c/2 - rint(c/2)|.
#pragma unroll
        for (int k = 0; k < 8; k++) {
            // (rint from the 1.5 * 2^23 trick: |c/2| < 2^22 here; no FRND on the slow pipe)
            const float t = cov[k] * 0.5f;
            const float nearest = __fadd_rn(__fadd_rn(t, 12582912.0f), -12582912.0f);
            cov[k] = 2.0f * fabsf(t - nearest);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) cov[k] = 1.0f;
    }
}


// ---- paints evaluated per pixel (GENERAL variant): colour textures with their filters
// (shaders/tile_fragment.inc.glsl:91-359: text, radial gradient, blur, colour matrix, none), SrcIn combine, and the
// blend modes (composite(), :414-535, + the blend states of renderer/src/gpu/blend.rs:43-163).
// `x`, `y` are texel coordinates (texel centres at integers, rows top-down). Sampler state = the batch's
// TextureSamplingFlags (gpu/src/lib.rs:521-528): LINEAR + CLAMP_TO_EDGE unless REPEAT_U / REPEAT_V / NEAREST_* say otherwise.
constexpr uint32_t SAMPLE_REPEAT_U = 0x1, SAMPLE_REPEAT_V = 0x2, SAMPLE_NEAREST_MIN = 0x4, SAMPLE_NEAREST_MAG = 0x8;

__device__ __forceinline__ int wrap_texel(int i, int n, bool repeat) {
    if (repeat) {
        i %= n;
        return i < 0 ? i + n : i;
    }
    return min(max(i, 0), n - 1);
}

__device__ __forceinline__ float4 sample_texture(const ColorTexture &t, float x, float y) {
    const bool rep_u = (t.sampling_flags & SAMPLE_REPEAT_U) != 0, rep_v = (t.sampling_flags & SAMPLE_REPEAT_V) != 0;
    auto texel = [&](int xx, int yy) {
        return unpack_rgba8(__ldg(reinterpret_cast<const uint32_t *>(t.pixels + (size_t)yy * t.pitch + (size_t)xx * 4)));
    };
    if (t.sampling_flags & (SAMPLE_NEAREST_MIN | SAMPLE_NEAREST_MAG)) // (the paint sets both or neither, paint.rs:560-563)
        return texel(wrap_texel((int)floorf(x + 0.5f), t.width, rep_u), wrap_texel((int)floorf(y + 0.5f), t.height, rep_v));
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    const int x0 = wrap_texel((int)fx, t.width, rep_u), x1 = wrap_texel((int)fx + 1, t.width, rep_u);
    const int y0 = wrap_texel((int)fy, t.height, rep_v), y1 = wrap_texel((int)fy + 1, t.height, rep_v);
    const float4 c00 = texel(x0, y0);
    if (ax == 0.0f && ay == 0.0f) return c00; // a texel centre: the usual case (render targets sampled pixel for pixel)
    const float4 c10 = texel(x1, y0), c01 = texel(x0, y1), c11 = texel(x1, y1);
    auto mix = [](float p, float q, float w) { return p + (q - p) * w; };
    return make_float4(mix(mix(c00.x, c10.x, ax), mix(c01.x, c11.x, ax), ay), mix(mix(c00.y, c10.y, ax), mix(c01.y, c11.y, ax), ay),
                       mix(mix(c00.z, c10.z, ax), mix(c01.z, c11.z, ax), ay), mix(mix(c00.w, c10.w, ax), mix(c01.w, c11.w, ax), ay));
}

// texture(colorTexture, uv): normalised coordinates -> texels (render targets are addressed bottom-up).
__device__ __forceinline__ float4 sample_uv(const ColorTexture &t, float u, float v) {
    const float x = u * (float)t.width - 0.5f;
    const float y = t.bottom_up ? ((float)t.height - 0.5f - v * (float)t.height) : (v * (float)t.height - 0.5f);
    return sample_texture(t, x, y);
}

// texture(gammaLUT, vec2(alpha, 1 - bg)).r: 256 x 8 L8, LINEAR, CLAMP_TO_EDGE (filterTextGammaCorrectChannel, :122-124).
__device__ __forceinline__ float sample_gamma(const uint8_t *__restrict__ lut, float u, float v) {
    const float x = u * 256.0f - 0.5f, y = v * 8.0f - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    const int x0 = min(max((int)fx, 0), 255), x1 = min(max((int)fx + 1, 0), 255);
    const int y0 = min(max((int)fy, 0), 7), y1 = min(max((int)fy + 1, 0), 7);
    const float s = 1.0f / 255.0f;
    const float top = (float)__ldg(lut + y0 * 256 + x0) * s * (1.0f - ax) + (float)__ldg(lut + y0 * 256 + x1) * s * ax;
    const float bottom = (float)__ldg(lut + y1 * 256 + x0) * s * (1.0f - ax) + (float)__ldg(lut + y1 * 256 + x1) * s * ax;
    return top * (1.0f - ay) + bottom * ay;
}

// filterText (tile_fragment.inc.glsl:91-166): nine taps one texel apart (onePixel = 1 / colorTextureSize.x), red
// channel only. p0 = kernel, p1 = bg, p2 = fg (w: gamma correction).
__device__ __forceinline__ float4 filter_text(const PaintTexture &p, const ColorTexture &t, const uint8_t *gamma_lut, float u, float v) {
    const float x = u * (float)t.width - 0.5f;
    const float y = t.bottom_up ? ((float)t.height - 0.5f - v * (float)t.height) : (v * (float)t.height - 0.5f);
    float3 alpha;
    if (p.p0.w == 0.0f) {
        const float r = sample_texture(t, x, y).x;
        alpha = make_float3(r, r, r);
    } else {
        const bool wide = p.p0.x > 0.0f;
        float tap[9];
#pragma unroll
        for (int k = 0; k < 9; k++) tap[k] = ((k == 0 || k == 8) && !wide) ? 0.0f : sample_texture(t, x + (float)(k - 4), y).x;
        // filterTextConvolve7Tap(alpha0, alpha1, kernel) = dot(alpha0, kernel) + dot(alpha1, kernel.zyx)
        auto convolve = [&](int first) {
            return (((tap[first] * p.p0.x + tap[first + 1] * p.p0.y) + tap[first + 2] * p.p0.z) + tap[first + 3] * p.p0.w) +
                   ((tap[first + 4] * p.p0.z + tap[first + 5] * p.p0.y) + tap[first + 6] * p.p0.x);
        };
        alpha = make_float3(convolve(0), convolve(1), convolve(2));
    }
    if (p.p2.w != 0.0f && gamma_lut) {
        alpha.x = sample_gamma(gamma_lut, alpha.x, 1.0f - p.p1.x);
        alpha.y = sample_gamma(gamma_lut, alpha.y, 1.0f - p.p1.y);
        alpha.z = sample_gamma(gamma_lut, alpha.z, 1.0f - p.p1.z);
    }
    // vec4(mix(bgColor, fgColor, alpha), 1.0)
    return make_float4(p.p1.x + (p.p2.x - p.p1.x) * alpha.x, p.p1.y + (p.p2.y - p.p1.y) * alpha.y,
                       p.p1.z + (p.p2.z - p.p1.z) * alpha.z, 1.0f);
}

// filterRadialGradient (tile_fragment.inc.glsl:274-300): p0 = line from, line vector; p1 = radii, uv origin.
__device__ __forceinline__ float4 filter_radial_gradient(const PaintTexture &p, const ColorTexture &t, float u, float v) {
    const float dpx = u - p.p0.x, dpy = v - p.p0.y, dcx = p.p0.z, dcy = p.p0.w;
    const float dr = p.p1.y - p.p1.x;
    const float a = (dcx * dcx + dcy * dcy) - dr * dr;
    const float b = (dpx * dcx + dpy * dcy) + p.p1.x * dr;
    const float c = (dpx * dpx + dpy * dpy) - p.p1.x * p.p1.x;
    const float discrim = b * b - a * c;
    if (discrim == 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float root = sqrtf(discrim);
    float t0 = (root + b) / a, t1 = (-root + b) / a;
    if (t0 > t1) {
        const float swap = t0;
        t0 = t1, t1 = swap;
    }
    const float tt = t0 >= 0.0f ? t0 : t1;
    return sample_uv(t, p.p1.z + tt, p.p1.w);
}

// filterBlur (tile_fragment.inc.glsl:302-339): p0 = direction, support; p1 = Gaussian coefficients.
__device__ __forceinline__ float4 filter_blur(const PaintTexture &p, const ColorTexture &t, float u, float v) {
    const float ox = p.p0.x / (float)t.width, oy = p.p0.y / (float)t.height;
    const int support = (int)p.p0.z;
    float gx = p.p1.x, gy = p.p1.y;
    const float gz = p.p1.z;
    float sum = gx;
    float4 color = sample_uv(t, u, v);
    color = make_float4(color.x * gx, color.y * gx, color.z * gx, color.w * gx);
    gx *= gy, gy *= gz;
    for (int i = 1; i <= support; i += 2) {
        float partial = gx;
        gx *= gy, gy *= gz;
        partial += gx;
        const float off = (float)i + gx / partial;
        const float4 lo = sample_uv(t, u - ox * off, v - oy * off), hi = sample_uv(t, u + ox * off, v + oy * off);
        color = make_float4(color.x + (lo.x + hi.x) * partial, color.y + (lo.y + hi.y) * partial,
                            color.z + (lo.z + hi.z) * partial, color.w + (lo.w + hi.w) * partial);
        sum += 2.0f * partial;
        gx *= gy, gy *= gz;
    }
    return make_float4(color.x / sum, color.y / sum, color.z / sum, color.w / sum);
}

// filterColor + combineColor0 (SrcIn) for the pixel whose centre is (fx, fy): the paint's colour, NOT premultiplied
// (tile_fragment.inc.glsl:81-89,361-412,583-603). A paint without a colour texture is its base colour.
__device__ __forceinline__ float4 paint_color(const PaintTexture &p, const ColorTexture &t, const uint8_t *gamma_lut,
                                              float fx, float fy) {
    if (!(p.flags & PAINT_HAS_TEXTURE) || t.pixels == nullptr) return p.base;
    const float u = p.m00 * fx + p.m01 * fy + p.tx, v = p.m10 * fx + p.m11 * fy + p.ty; // computeTileVaryings
    float4 c;
    switch (p.filter_kind) {
    case 1: c = filter_radial_gradient(p, t, u, v); break;                 // PF_FILTER_RADIAL_GRADIENT
    case PF_FILTER_TEXT_KIND: c = filter_text(p, t, gamma_lut, u, v); break;
    case 3: c = filter_blur(p, t, u, v); break;                            // PF_FILTER_BLUR
    case 4: {                                                              // PF_FILTER_COLOR_MATRIX: matrix * colour + p4
        const float4 q = sample_uv(t, u, v);
        c = make_float4(p.p0.x * q.x + p.p1.x * q.y + p.p2.x * q.z + p.p3.x * q.w + p.p4.x,
                        p.p0.y * q.x + p.p1.y * q.y + p.p2.y * q.z + p.p3.y * q.w + p.p4.y,
                        p.p0.z * q.x + p.p1.z * q.y + p.p2.z * q.z + p.p3.z * q.w + p.p4.z,
                        p.p0.w * q.x + p.p1.w * q.y + p.p2.w * q.z + p.p3.w * q.w + p.p4.w);
        break;
    }
    default: c = sample_uv(t, u, v); break;                                // filterNone
    }
    // combineColor0 (tile_fragment.inc.glsl:81-89; dest = the base colour, src = the filtered texture colour):
    // SrcIn (src.rgb, src.a * dest.a), DestIn (dest.rgb, src.a * dest.a)
    if (p.flags & PAINT_COMBINE_DEST_IN) return make_float4(p.base.x, p.base.y, p.base.z, c.w * p.base.w);
    return make_float4(c.x, c.y, c.z, c.w * p.base.w);
}

// ---- blend modes. Values of PF_BLEND_MODE_* (include/pf_cuda.h).
__device__ __forceinline__ float composite_divide(float num, float denom) { return denom != 0.0f ? num / denom : 0.0f; }
__device__ __forceinline__ float color_dodge1(float d, float s) { return d == 0.0f ? 0.0f : (s == 1.0f ? 1.0f : d / (1.0f - s)); }
__device__ __forceinline__ float screen1(float d, float s) { return d + s - d * s; }
__device__ __forceinline__ float hard_light1(float d, float s) { return s <= 0.5f ? d * 2.0f * s : screen1(d, 2.0f * s - 1.0f); }
__device__ __forceinline__ float soft_light1(float d, float s) {
    const float darkened = d <= 0.25f ? ((16.0f * d - 12.0f) * d + 4.0f) * d : sqrtf(d);
    const float factor = s <= 0.5f ? d * (1.0f - d) : darkened - d;
    return d + (s * 2.0f - 1.0f) * factor;
}
__device__ __forceinline__ float3 rgb_to_hsl(float3 c) { // compositeRGBToHSL (:443-453)
    const float v = fmaxf(fmaxf(c.x, c.y), c.z), x_min = fminf(fminf(c.x, c.y), c.z);
    const float ch = v - x_min, l = x_min + (v - x_min) * 0.5f;
    float3 terms = c.x == v ? make_float3(0.0f, c.y, c.z) : (c.y == v ? make_float3(2.0f, c.z, c.x) : make_float3(4.0f, c.x, c.y));
    const float h = 1.0471975511965976f * composite_divide(terms.x * ch + terms.y - terms.z, ch); // FRAC_PI_3
    return make_float3(h, composite_divide(ch, v), l);
}
__device__ __forceinline__ float3 hsl_to_rgb(float3 hsl) { // compositeHSLToRGB (:436-440)
    const float a = hsl.y * fminf(hsl.z, 1.0f - hsl.z);
    const float hk = hsl.x * 1.9098593171027440f; // FRAC_6_PI
    auto channel = [&](float n) {
        float k = n + hk;
        k = k - 12.0f * floorf(k / 12.0f); // mod(k, 12)
        return hsl.z - fminf(fmaxf(fminf(k - 3.0f, 9.0f - k), -1.0f), 1.0f) * a;
    };
    return make_float3(channel(0.0f), channel(8.0f), channel(4.0f));
}
// compositeRGB (:488-523), for the blend modes the shader evaluates itself.
__device__ __forceinline__ float3 composite_rgb(float3 d, float3 s, uint32_t mode) {
    switch (mode) {
    case 14: return make_float3(d.x * s.x, d.y * s.y, d.z * s.z);                                        // Multiply
    case 15: return make_float3(screen1(d.x, s.x), screen1(d.y, s.y), screen1(d.z, s.z));                // Screen
    case 17: return make_float3(hard_light1(s.x, d.x), hard_light1(s.y, d.y), hard_light1(s.z, d.z));    // Overlay
    case 12: return make_float3(fminf(d.x, s.x), fminf(d.y, s.y), fminf(d.z, s.z));                      // Darken
    case 13: return make_float3(fmaxf(d.x, s.x), fmaxf(d.y, s.y), fmaxf(d.z, s.z));                      // Lighten
    case 18: return make_float3(color_dodge1(d.x, s.x), color_dodge1(d.y, s.y), color_dodge1(d.z, s.z)); // ColorDodge
    case 19:                                                                                              // ColorBurn
        return make_float3(1.0f - color_dodge1(1.0f - d.x, 1.0f - s.x), 1.0f - color_dodge1(1.0f - d.y, 1.0f - s.y),
                           1.0f - color_dodge1(1.0f - d.z, 1.0f - s.z));
    case 16: return make_float3(hard_light1(d.x, s.x), hard_light1(d.y, s.y), hard_light1(d.z, s.z));    // HardLight
    case 20: return make_float3(soft_light1(d.x, s.x), soft_light1(d.y, s.y), soft_light1(d.z, s.z));    // SoftLight
    case 21: return make_float3(fabsf(d.x - s.x), fabsf(d.y - s.y), fabsf(d.z - s.z));                   // Difference
    case 22: return make_float3(d.x + s.x - 2.0f * d.x * s.x, d.y + s.y - 2.0f * d.y * s.y, d.z + s.z - 2.0f * d.z * s.z); // Exclusion
    case 23: case 24: case 25: case 26: {                                                                // Hue, Saturation, Color, Luminosity
        const float3 dh = rgb_to_hsl(d), sh = rgb_to_hsl(s);
        const float3 pick = mode == 23 ? make_float3(sh.x, dh.y, dh.z)
                          : mode == 24 ? make_float3(dh.x, sh.y, dh.z)
                          : mode == 25 ? make_float3(sh.x, sh.y, dh.z)
                                       : make_float3(dh.x, dh.y, sh.z);
        return hsl_to_rgb(pick);
    }
    default: return s;
    }
}

// One pixel of one path: dest (premultiplied) under the paint's colour c (not premultiplied) with mask alpha m.
// SrcOver: dest * (1 - a) + (rgb * a, a). The other Porter-Duff modes: the blend factors of gpu/blend.rs:43-143 on the
// premultiplied source. The modes the shader evaluates itself (Darken .. Luminosity): composite() reads the
// destination, mixes, forces alpha to 1 and the result replaces the pixel (blending disabled, blend.rs:144-163).
__device__ __forceinline__ float4 blend_pixel(float4 d, float4 c, float m, uint32_t mode) {
    const float sa = c.w * m;
    if (mode >= 12u) { // PF_BLEND_MODE_DARKEN ..
        const float3 b = composite_rgb(make_float3(d.x, d.y, d.z), make_float3(c.x, c.y, c.z), mode);
        const float k0 = sa * (1.0f - d.w), k1 = sa * d.w, k2 = 1.0f - sa;
        // (dodge and burn leave [0, 1]; the render target is UNORM, and so is what the next path reads)
        return make_float4(__saturatef(k0 * c.x + k1 * b.x + k2 * d.x), __saturatef(k0 * c.y + k1 * b.y + k2 * d.y),
                           __saturatef(k0 * c.z + k1 * b.z + k2 * d.z), 1.0f);
    }
    float sf, df; // source / destination factors (the same for colour and alpha in every mode)
    switch (mode) {
    // (the destructive modes, effects.rs:222-235: what they do to pixels the mask leaves out — sa = 0 — is part of
    // the mode; a tile the path does not reach is never drawn, builder.rs:1014-1016)
    case 0: sf = 0.0f, df = 0.0f; break;              // Clear
    case 1: sf = 1.0f, df = 0.0f; break;              // Copy (blending disabled, blend.rs:144)
    case 2: sf = d.w, df = 0.0f; break;               // SrcIn
    case 3: sf = 1.0f - d.w, df = 0.0f; break;        // SrcOut
    case 6: sf = 0.0f, df = sa; break;                // DestIn
    case 9: sf = 1.0f - d.w, df = sa; break;          // DestAtop
    case 8: sf = 1.0f - d.w, df = 1.0f; break;        // DestOver
    case 7: sf = 0.0f, df = 1.0f - sa; break;         // DestOut
    case 5: sf = d.w, df = 1.0f - sa; break;          // SrcAtop
    case 10: sf = 1.0f - d.w, df = 1.0f - sa; break;  // Xor
    case 11: sf = 1.0f, df = 1.0f; break;             // Lighter
    default: sf = 1.0f, df = 1.0f - sa; break;        // SrcOver
    }
    const float k = sa * sf;
    float4 out = make_float4(fmaf(d.x, df, c.x * k), fmaf(d.y, df, c.y * k), fmaf(d.z, df, c.z * k), fmaf(d.w, df, k));
    // The render target is UNORM: Lighter adds, and a colour-matrix filter can hand out colours outside [0, 1].
    return make_float4(__saturatef(out.x), __saturatef(out.y), __saturatef(out.z), __saturatef(out.w));
}

template <bool LOAD_DEST, bool GENERAL>
__global__ void __launch_bounds__(32 * TILE_WARPS, GENERAL ? 4 : PF_TILE_MIN_BLOCKS) k_tile_alpha(CompositeArgs a) {
    chain_wait(); // (programmatic dependent launch: nothing is touched before the stage before this one has completed)
    __shared__ TileWarpShared<GENERAL> sh_all[TILE_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TileWarpShared<GENERAL> &sh = sh_all[warp];
    const uint32_t n_queue = *a.queue_count; // written by k_tile_solid
    if (a.export_alpha_count && blockIdx.x == 0 && threadIdx.x == 0) *a.export_alpha_count = n_queue;
    const int x = lane & 15, half = lane >> 4;
    const float xf = (float)x;
    const float u_off = 0.5f - 0.5f * (float)half; // (8 - 8 * half) / 16
    const bool simple_store = a.n_dest == 1 && a.export_blocks == nullptr && (a.dest_align_mask & 15u) == 0;

    // Persistent warps: tiles differ wildly in depth, so every warp pulls its next tile from a global
    // counter. The pull is software-pipelined two deep: while tile i is composited, the counter increment
    // for tile i + 2 and the queue slot + list header of tile i + 1 are in flight.
    auto claim = [&]() -> uint32_t {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(a.work_counter, 1u);
        return c; // lane 0 only; broadcast when it is needed, not before
    };
    auto fetch = [&](uint32_t k, uint32_t &work, uint32_t &count, uint32_t &start) {
        work = count = start = 0;
        if (k < n_queue) {
            work = __ldg(a.queue + k);
            const uint2 hdr = __ldg(a.queue_hdr + k);
            count = hdr.x, start = hdr.y;
        }
    };
    uint32_t pending = claim();
    uint32_t k_cur = __shfl_sync(0xffffffffu, pending, 0);
    pending = claim();
    uint32_t work, n, e0;
    fetch(k_cur, work, n, e0);
    for (;;) {
        if (k_cur >= n_queue) return;
        const uint32_t k_next = __shfl_sync(0xffffffffu, pending, 0);
        pending = claim();
        uint32_t work_next, n_next, e0_next;
        fetch(k_next, work_next, n_next, e0_next);

        const int ty = a.tile_y0 + (int)(work >> 16), tx = a.fb.min_x + (int)(work & 0xffffu);
        const int px = tx * 16 + x, py0 = ty * 16 + half * 8;

        // ---- bring the run into draw order: rank sort by tile index (tiles are allocated path by path, so
        // ascending tile index is draw order; replaces the insertion sort of shaders/d3d11/sort.cs.glsl).
        const bool in_smem = n <= (uint32_t)ENTRY_CAP;
        if (in_smem) {
            uint4 raw = make_uint4(0, 0, 0, 0xffffffffu);
            float4 paint = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if ((uint32_t)lane < n) {
                raw = __ldg(reinterpret_cast<const uint4 *>(a.entries + e0 + lane));
                paint = __ldg(&a.entries[e0 + lane].color);
#if PF_PREFETCH
                if (raw.y & 0x00ffffffu) {
                    const uint32_t first = raw.x - (raw.y & 0x00ffffffu);
                    prefetch_l2(a.fills + first);
                }
#endif
            }
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; j++) rank += __shfl_sync(0xffffffffu, raw.w, j) < raw.w ? 1u : 0u;
            if ((uint32_t)lane < n) {
                sh.entry[rank] = raw;
                sh.paint[rank] = paint;
                if (GENERAL && a.entry_clip) sh.clip[rank] = __ldg(a.entry_clip + e0 + lane);
            }
            __syncwarp();
        }

        // Pixel state: while `expanded` is false every pixel is `o`; afterwards pixel k of this lane is
        // sh.dst[k * 32 + lane] * s + o (solid entries only update s and o).
        bool expanded = false;
        float s = 1.0f;
        Px o = px_from(a.clear_color);
        if (LOAD_DEST) {
            // Rows / columns outside the image read as the clear colour (they are never stored).
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float4 d = a.clear_color;
                const int py = py0 + k;
                if (px >= 0 && px < a.dest_w && py >= 0 && py < a.dest_h)
                    d = unpack_rgba8(*reinterpret_cast<const uint32_t *>(a.dest + (size_t)py * a.dest_pitch + (size_t)px * 4));
                sh.dst[k * 32 + lane] = d;
            }
            expanded = true;
            o = px_from(make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        }

        uint32_t next_key = 0; // selection path cursor: smallest tile index not yet drawn
        for (uint32_t ei = 0; ei < n; ei++) {
            uint4 raw;
            float4 paint;
            uint2 clip_entry = make_uint2(0, 0);
            if (in_smem) {
                raw = sh.entry[ei];
                paint = sh.paint[ei];
                if (GENERAL && a.entry_clip) clip_entry = sh.clip[ei];
            } else {
                // Very deep lists: select the next entry in draw order by a min-scan.
                uint32_t best = 0xffffffffu, best_i = 0;
                for (uint32_t i = lane; i < n; i += 32) {
                    const uint32_t key = __ldg(&a.entries[e0 + i].tile_index);
                    if (key >= next_key && key < best) best = key, best_i = i;
                }
                for (int d = 16; d > 0; d >>= 1) {
                    const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, d), oi = __shfl_xor_sync(0xffffffffu, best_i, d);
                    if (ob < best) best = ob, best_i = oi;
                }
                next_key = best + 1;
                raw = __ldg(reinterpret_cast<const uint4 *>(a.entries + e0 + best_i));
                paint = __ldg(&a.entries[e0 + best_i].color);
                if (GENERAL && a.entry_clip) clip_entry = __ldg(a.entry_clip + e0 + best_i);
            }
            const uint32_t fill_end = raw.x, count = raw.y & 0x00ffffffu;
            const float backdrop = (float)(int)(int8_t)(raw.y >> 24);
            const uint32_t ctrl = (raw.z >> 16) & 0xffu;
            const Px pp = px_from(paint);
            const f32x2 neg_w = pack2(-paint.w, -paint.w);
            const bool clipped = GENERAL && (raw.z & ENTRY_HAS_CLIP);
            const bool textured = GENERAL && (raw.z & ENTRY_TEXTURED);

            if (count == 0 && !clipped && !textured) {
                // Solid tile: coverage = backdrop for every pixel (tile_fragment.inc.glsl:548) — the same
                // blend for all 256 pixels, folded into the affine map.
                const float m = mask_alpha(backdrop, ctrl);
                s = fmaf(-paint.w * m, s, s);
                over(o, pp, neg_w, m);
                continue;
            }

            float cov[8];
            if (!clipped && count == 0) { // a solid tile of a textured paint: the same mask for every pixel
                const float m = mask_alpha(backdrop, ctrl);
#pragma unroll
                for (int k = 0; k < 8; k++) cov[k] = m;
            } else if (!clipped) {
                fill_coverage<GENERAL>(sh, a.fills, fill_end, count, backdrop, xf, u_off, a.area_lut, lane, cov);
                rule_of(cov, ctrl);
            } else {
                // A tile of a clipped path that meets an alpha tile of its clip path (tiler.rs:114-156). D3D9
                // combines the masks as min(|draw + backdrop|, |clip + backdrop|) with the draw tile's
                // backdrop then zeroed (tile_clip_combine.fs.glsl:28-31); a solid draw tile simply takes
                // over the clip tile's mask and backdrop.
                const bool replace = (raw.z & ENTRY_CLIP_REPLACE) != 0;
                const uint32_t clip_count = clip_entry.y & 0x00ffffffu;
                const float clip_backdrop = (float)(int)(int8_t)(clip_entry.y >> 24);
                fill_coverage<GENERAL>(sh, a.clip_fills, clip_entry.x, clip_count, clip_backdrop, xf, u_off, a.area_lut, lane, cov);
                if (!replace) {
                    float cov_draw[8];
                    fill_coverage<GENERAL>(sh, a.fills, fill_end, count, backdrop, xf, u_off, a.area_lut, lane, cov_draw);
#pragma unroll
                    for (int k = 0; k < 8; k++) cov[k] = fminf(fabsf(cov_draw[k]), fabsf(cov[k]));
                }
                rule_of(cov, ctrl);
            }

            // calculateColor (tile_fragment.inc.glsl:560-614), SrcOver — per pixel.
            const f32x2 ss = pack2(s, s);
            if (textured) {
                // The paint samples the batch's colour texture: a colour per pixel (filterColor + combineColor0).
                const PaintTexture pt = a.paint_textures[raw.z & 0xffffu];
                const float4 o4 = px_to(o);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    float4 d = o4;
                    if (expanded) {
                        const float4 v = sh.dst[k * 32 + lane];
                        d = make_float4(fmaf(v.x, s, o4.x), fmaf(v.y, s, o4.y), fmaf(v.z, s, o4.z), fmaf(v.w, s, o4.w));
                    }
                    const float4 c = paint_color(pt, a.color_texture, a.gamma_lut, (float)px + 0.5f, (float)(py0 + k) + 0.5f);
                    // color.a *= maskAlpha; composite; color.rgb *= color.a; blend (tile_fragment.inc.glsl:583-614)
                    d = blend_pixel(d, c, cov[k], pt.blend_mode);
                    sh.dst[k * 32 + lane] = d;
                }
                expanded = true;
            } else if (expanded) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float4 v = sh.dst[k * 32 + lane];
                    Px d{fma2(pack2(v.x, v.y), ss, o.rg), fma2(pack2(v.z, v.w), ss, o.ba)};
                    over(d, pp, neg_w, cov[k]);
                    sh.dst[k * 32 + lane] = px_to(d);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    Px d = o;
                    over(d, pp, neg_w, cov[k]);
                    sh.dst[k * 32 + lane] = px_to(d);
                }
                expanded = true;
            }
            s = 1.0f;
            o = px_from(make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        }

        // ---- store. The tile is transposed through shared memory so that it leaves as 128-bit stores (two
        // per lane) instead of eight 32-bit ones; when the frame is strip-partitioned over several GPUs the
        // same stores also go straight into every peer's copy of the frame over NVLink.
        uint32_t pk[8];
        if (expanded) {
            // (v * s + o) * 255 + 2^23 = v * (255 s) + (255 o + 2^23): the affine map and the scaling in one FMA
            const f32x2 scale = pack2(255.0f, 255.0f), magic = pack2(8388608.0f, 8388608.0f);
            const f32x2 s255 = pack2(s * 255.0f, s * 255.0f);
            const f32x2 o_rg = fma2(o.rg, scale, magic), o_ba = fma2(o.ba, scale, magic);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 v = sh.dst[k * 32 + lane];
                float r, g, b, al;
                unpack2(fma2(pack2(v.x, v.y), s255, o_rg), r, g);
                unpack2(fma2(pack2(v.z, v.w), s255, o_ba), b, al);
                const uint32_t lo = __byte_perm(__float_as_uint(r), __float_as_uint(g), 0x0040);
                const uint32_t hi = __byte_perm(__float_as_uint(b), __float_as_uint(al), 0x0040);
                pk[k] = __byte_perm(lo, hi, 0x5410);
            }
        } else {
            const uint32_t packed = pack_rgba8(o);
#pragma unroll
            for (int k = 0; k < 8; k++) pk[k] = packed;
        }
        const bool inside = tx >= 0 && ty >= 0 && tx * 16 + 16 <= a.dest_w && ty * 16 + 16 <= a.dest_h;
        if (PF_SIMPLE_STORE && simple_store && inside) {
            // One local, aligned image, nothing exported: no loop over destinations, no per-destination checks.
#pragma unroll
            for (int k = 0; k < 8; k++) sh.stage[(half * 8 + k) * 16 + half * 16 + x] = pk[k];
            __syncwarp();
            const int row0 = lane >> 2, quarter = lane & 3;
            const uint4 v0 = *reinterpret_cast<const uint4 *>(&sh.stage[row0 * 16 + quarter * 4]);
            const uint4 v1 = *reinterpret_cast<const uint4 *>(&sh.stage[(row0 + 8) * 16 + 16 + quarter * 4]);
            uint8_t *out = a.dest + (size_t)(ty * 16 + row0) * a.dest_pitch + (size_t)(tx * 64 + quarter * 16);
            *reinterpret_cast<uint4 *>(out) = v0;
            *reinterpret_cast<uint4 *>(out + 8 * a.dest_pitch) = v1;
        } else {
        const bool vector_store = inside && (a.dest_align_mask & 15u) == 0;
        if (vector_store || a.export_blocks) {
#pragma unroll
            for (int k = 0; k < 8; k++) sh.stage[(half * 8 + k) * 16 + half * 16 + x] = pk[k];
            __syncwarp();
            // chunk c = lane (rows 0..7) and lane + 32 (rows 8..15): row c / 4, 16-byte quarter c % 4
            const int row0 = lane >> 2, quarter = lane & 3;
            const uint4 v0 = *reinterpret_cast<const uint4 *>(&sh.stage[row0 * 16 + quarter * 4]);
            const uint4 v1 = *reinterpret_cast<const uint4 *>(&sh.stage[(row0 + 8) * 16 + 16 + quarter * 4]);
            if (a.export_blocks) { // slot k_cur: the tile as one contiguous 1 KB block for the other ranks
                uint4 *block = reinterpret_cast<uint4 *>(a.export_blocks + (size_t)k_cur * 1024u);
                block[lane] = v0;
                block[lane + 32] = v1;
            }
            if (vector_store) {
                const size_t off0 = (size_t)(ty * 16 + row0) * a.dest_pitch + (size_t)tx * 64 + (size_t)quarter * 16;
                const size_t off1 = off0 + 8 * a.dest_pitch;
                for (int d = 0; d < a.n_dest; d++) {
                    *reinterpret_cast<uint4 *>(a.dests[d] + off0) = v0;
                    *reinterpret_cast<uint4 *>(a.dests[d] + off1) = v1;
                }
            }
        }
        if (!vector_store && px >= 0 && px < a.dest_w) {
            for (int d = 0; d < a.n_dest; d++) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int py = py0 + k;
                    if (py >= 0 && py < a.dest_h)
                        *reinterpret_cast<uint32_t *>(a.dests[d] + (size_t)py * a.dest_pitch + (size_t)px * 4) = pk[k];
                }
            }
        }
        }
        __syncwarp(); // the next tile reuses this warp's shared-memory slots
#if PF_PREFETCH
        if ((uint32_t)lane < n_next && n_next <= (uint32_t)ENTRY_CAP) prefetch_l2(a.entries + e0_next + lane);
#endif
        k_cur = k_next, work = work_next, n = n_next, e0 = e0_next;
    }
}

// Resident blocks of the persistent kernel per device (one process may drive several GPUs).
struct Residency {
    int sm_count = 0;
    int blocks[2][2] = {{0, 0}, {0, 0}}; // [load_dest][has_clip]
};
Residency &residency_of_current_device() {
    static Residency table[64];
    int dev = 0;
    PF_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) throw Error(2, "device ordinal out of range");
    Residency &r = table[dev];
    if (r.sm_count == 0) {
        int sm = 0;
        PF_CUDA_CHECK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
        const int threads = 32 * TILE_WARPS;
        PF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.blocks[0][0], k_tile_alpha<false, false>, threads, 0));
        PF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.blocks[1][0], k_tile_alpha<true, false>, threads, 0));
        PF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.blocks[0][1], k_tile_alpha<false, true>, threads, 0));
        PF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.blocks[1][1], k_tile_alpha<true, true>, threads, 0));
        r.sm_count = sm;
    }
    return r;
}

} // namespace

int launch_composite(const CompositeArgs &args, cudaStream_t stream) {
    const CompositeArgs &a = args;
    const bool has_clip = a.entry_clip != nullptr || a.paint_textures != nullptr; // the GENERAL kernel variants
    const int fb_w = a.fb.max_x - a.fb.min_x;
    const int rows = a.tile_y1 - a.tile_y0;
    if (fb_w <= 0 || rows <= 0) return 0;
    if (fb_w > 65535 || rows > 65535) throw Error(2, "framebuffer larger than 65535 tiles on a side");
    // The caller has zeroed a.work_counter and a.queue_count (the tile-list kernel may since have parked the
    // work counter past the end).
    const uint64_t segments = (uint64_t)((fb_w + 31) / 32) * (uint64_t)rows;
    const unsigned solid_grid = (unsigned)((segments + 3) / 4); // 4 warps per block
    if (a.load_dest)
        launch_chained(k_tile_solid<true>, solid_grid, 128, stream, a);
    else
        launch_chained(k_tile_solid<false>, solid_grid, 128, stream, a);
    PF_CUDA_CHECK(cudaGetLastError());
    if (!a.entries) return 1; // a frame without batches: every list is empty, nothing was queued

    // One resident wave of persistent warps, or fewer when the whole frame has fewer tiles.
    const Residency &res = residency_of_current_device();
    const uint64_t n_work = (uint64_t)fb_w * (uint64_t)rows;
    const uint64_t want = (n_work + TILE_WARPS - 1) / TILE_WARPS;
    const uint64_t resident = (uint64_t)res.sm_count * (uint64_t)res.blocks[a.load_dest ? 1 : 0][has_clip ? 1 : 0];
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min(want, resident));
    const int threads = 32 * TILE_WARPS;
    if (has_clip) {
        if (a.load_dest)
            launch_chained(k_tile_alpha<true, true>, grid, threads, stream, a);
        else
            launch_chained(k_tile_alpha<false, true>, grid, threads, stream, a);
    } else if (a.load_dest) {
        launch_chained(k_tile_alpha<true, false>, grid, threads, stream, a);
    } else {
        launch_chained(k_tile_alpha<false, false>, grid, threads, stream, a);
    }
    PF_CUDA_CHECK(cudaGetLastError());
    return 2;
}

// ---------------------------------------------------------------------------------------------
// k_push_export / k_pull_tiles — frame assembly across GPUs from compact exports. The fill + tile kernels leave a
// compact export of their strip next to the frame (4 bytes per single-colour tile, one contiguous 1 KB block per other
// tile); k_push_export copies the used part of it into a receive slot in every other rank's memory over NVLink
// (coalesced 16-byte posted writes, the source read once for all peers), and after a barrier k_pull_tiles expands what
// arrived — all of it local memory by then — into the rank's own copy of the frame. Both run on the gather stream
// beside the next frame's stages. Compared with an all-gather of the finished strips a fraction of the bytes crosses
// the links (tiger frames: a few per cent; random100k: under half). Measured alternatives (B200 x 8): peers READING
// the exports over NVLink (340 GB/s, latency-bound), and the compositing kernel storing the blocks into the peers
// itself (480 GB/s, but serialised with the rank's own compute instead of overlapping the next frame).
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(128) k_pull_tiles(PullArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int fb_w = a.fb.max_x - a.fb.min_x;
    const uint32_t segs = ((uint32_t)fb_w + 31u) >> 5;
    const bool aligned = (((uintptr_t)a.dest | (uintptr_t)a.dest_pitch) & 15u) == 0;
    for (int p = 0; p < a.n_peers; p++) {
        const PullPeer peer = a.peers[p];
        // ---- single-colour tiles, one 32-tile row segment per warp iteration (as k_tile_solid stores them)
        const uint32_t n_segments = segs * (uint32_t)(peer.tile_y1 - peer.tile_y0);
        for (uint32_t segment = warp_global; segment < n_segments; segment += n_warps) {
            const uint32_t paint_mask = __ldg(peer.solid_mask + segment);
            if (paint_mask == 0) continue;
            const uint32_t packed = __ldg(peer.solid_color + (size_t)segment * 32u + (uint32_t)lane);
            const uint32_t tile_row = segment / segs, seg = segment - tile_row * segs;
            const int ty = peer.tile_y0 + (int)tile_row;
            const int seg_tx = a.fb.min_x + (int)(seg * 32u);
            uint32_t chunk_color[4];
            bool chunk_on[4], chunk_full[4];
            int chunk_px[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = q * 8 + (lane >> 2);
                chunk_color[q] = __shfl_sync(0xffffffffu, packed, j);
                chunk_px[q] = (seg_tx + j) * 16 + (lane & 3) * 4;
                chunk_on[q] = ((paint_mask >> j) & 1u) && chunk_px[q] + 4 > 0 && chunk_px[q] < a.dest_w;
                chunk_full[q] = chunk_px[q] >= 0 && chunk_px[q] + 4 <= a.dest_w && aligned;
            }
            const int y_lo = max(0, ty * 16), y_hi = min(a.dest_h, ty * 16 + 16);
            uint8_t *row = a.dest + (ptrdiff_t)y_lo * (ptrdiff_t)a.dest_pitch;
            for (int y = y_lo; y < y_hi; y++, row += a.dest_pitch) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (!chunk_on[q]) continue;
                    if (chunk_full[q]) {
                        *reinterpret_cast<uint4 *>(row + (ptrdiff_t)chunk_px[q] * 4) =
                            make_uint4(chunk_color[q], chunk_color[q], chunk_color[q], chunk_color[q]);
                    } else {
                        for (int k = 0; k < 4; k++) {
                            const int px = chunk_px[q] + k;
                            if (px >= 0 && px < a.dest_w) *reinterpret_cast<uint32_t *>(row + (ptrdiff_t)px * 4) = chunk_color[q];
                        }
                    }
                }
            }
        }
        // ---- the other tiles, 32 per warp iteration: one coalesced read of their positions, then the 1 KB blocks
        // four at a time (eight 128-bit loads in flight per lane) and
        // two 128-bit stores per lane and tile into the local frame.
        const uint32_t n_alpha = __ldg(peer.alpha_count);
        for (uint32_t k0 = warp_global * 32u; k0 < n_alpha; k0 += n_warps * 32u) {
            const uint32_t m = min(32u, n_alpha - k0);
            const uint32_t my_work = (uint32_t)lane < m ? __ldg(peer.queue + k0 + lane) : 0u;
            const uint4 *blocks = reinterpret_cast<const uint4 *>(peer.blocks + (size_t)k0 * 1024u);
            for (uint32_t t0 = 0; t0 < m; t0 += 4) {
                uint4 v0[4], v1[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (t0 + i < m) {
                        v0[i] = __ldg(blocks + (size_t)(t0 + i) * 64u + lane);
                        v1[i] = __ldg(blocks + (size_t)(t0 + i) * 64u + lane + 32);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t work = __shfl_sync(0xffffffffu, my_work, (int)((t0 + i) & 31u));
                    if (t0 + i >= m) continue;
                    const int ty = peer.tile_y0 + (int)(work >> 16), tx = a.fb.min_x + (int)(work & 0xffffu);
                    const int row0 = lane >> 2, quarter = lane & 3;
                    const bool inside = tx >= 0 && ty >= 0 && tx * 16 + 16 <= a.dest_w && ty * 16 + 16 <= a.dest_h;
                    if (inside && aligned) {
                        uint8_t *o = a.dest + (size_t)(ty * 16 + row0) * a.dest_pitch + (size_t)tx * 64 + (size_t)quarter * 16;
                        *reinterpret_cast<uint4 *>(o) = v0[i];
                        *reinterpret_cast<uint4 *>(o + 8 * a.dest_pitch) = v1[i];
                    } else {
                        const uint32_t w0[4] = {v0[i].x, v0[i].y, v0[i].z, v0[i].w}, w1[4] = {v1[i].x, v1[i].y, v1[i].z, v1[i].w};
                        for (int c = 0; c < 4; c++) {
                            const int px = tx * 16 + quarter * 4 + c;
                            if (px < 0 || px >= a.dest_w) continue;
                            const int py_a = ty * 16 + row0, py_b = py_a + 8;
                            if (py_a >= 0 && py_a < a.dest_h) *reinterpret_cast<uint32_t *>(a.dest + (size_t)py_a * a.dest_pitch + (size_t)px * 4) = w0[c];
                            if (py_b >= 0 && py_b < a.dest_h) *reinterpret_cast<uint32_t *>(a.dest + (size_t)py_b * a.dest_pitch + (size_t)px * 4) = w1[c];
                        }
                    }
                }
            }
        }
    }
}

} // namespace

namespace {
__global__ void __launch_bounds__(256) k_push_export(PushArgs a) {
    const uint32_t n_alpha = *reinterpret_cast<const uint32_t *>(a.local);
    // the used ranges of the slot, in 16-byte units: {count}, queue, colours, masks, blocks
    const size_t begin[5] = {0, a.queue_off / 16, a.color_off / 16, a.mask_off / 16, a.blocks_off / 16};
    const size_t length[5] = {1, ((size_t)n_alpha * 4 + 15) / 16, ((size_t)a.segments * 128 + 15) / 16,
                              ((size_t)a.segments * 4 + 15) / 16, (size_t)n_alpha * 64};
    const uint4 *src = reinterpret_cast<const uint4 *>(a.local);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (int range = 0; range < 5; range++) {
        const uint4 *s = src + begin[range];
        size_t i = tid;
        for (; i + 3 * stride < length[range]; i += 4 * stride) { // four loads in flight, each stored to every peer
            const uint4 v0 = s[i], v1 = s[i + stride], v2 = s[i + 2 * stride], v3 = s[i + 3 * stride];
            for (int p = 0; p < a.n_remote; p++) {
                uint4 *d = reinterpret_cast<uint4 *>(a.remote[p]) + begin[range];
                d[i] = v0, d[i + stride] = v1, d[i + 2 * stride] = v2, d[i + 3 * stride] = v3;
            }
        }
        for (; i < length[range]; i += stride) {
            const uint4 v = s[i];
            for (int p = 0; p < a.n_remote; p++) (reinterpret_cast<uint4 *>(a.remote[p]) + begin[range])[i] = v;
        }
    }
}
} // namespace

int launch_push_export(const PushArgs &args, cudaStream_t stream) {
    if (args.n_remote <= 0) return 0;
    int dev = 0, sm = 0;
    PF_CUDA_CHECK(cudaGetDevice(&dev));
    PF_CUDA_CHECK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    // NVLink-bound: a block per SM keeps the links busy and leaves the SMs to the next frame's stages.
    k_push_export<<<(unsigned)sm, 256, 0, stream>>>(args);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

int launch_pull_tiles(const PullArgs &args, cudaStream_t stream) {
    if (args.n_peers <= 0) return 0;
    int dev = 0, sm = 0;
    PF_CUDA_CHECK(cudaGetDevice(&dev));
    PF_CUDA_CHECK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    // Latency-bound remote reads: many warps in flight, but not the whole GPU — the next frame's stages run beside it.
    k_pull_tiles<<<(unsigned)sm * 8u, 128, 0, stream>>>(args);
    PF_CUDA_CHECK(cudaGetLastError());
    return 1;
}

} // namespace pf
