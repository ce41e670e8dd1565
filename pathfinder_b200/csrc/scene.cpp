// pathfinder_b200/csrc/scene.cpp — host side above the C ABI: Scene, BuildOptions and the D3D11-level
// SceneBuilder, mirrored in C++ because this image has no Rust toolchain (SURVEY.md §0 finding 2).
//
// Mirrors (paths relative to the reference checkout):
//   Scene / DrawPath / ClipPath          renderer/src/scene.rs:36-226, 425-560
//   BuildOptions / RenderTransform       renderer/src/options.rs:50-181
//   SceneBuilder::build (D3D11 branch)   renderer/src/builder.rs:148-222
//   BuiltSegments::from_scene, add_path  renderer/src/builder.rs:777-841
//   prepare_draw_path_for_gpu_binning    renderer/src/builder.rs:1058-1095
//   TileBatchDataD3D11::push             renderer/src/builder.rs:653-759
//   Palette (solid colours only)         renderer/src/paint.rs:398-438, 641-659
// Geometry helpers keep the reference's operator order so tile rects match
// the CPU tiler's for scale + translate transforms; build without -ffast-math / FMA contraction.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/pf_cuda.h"
#include "common.cuh"
#include "dilate.h"

#include <cuda_runtime.h>

namespace pf {
thread_local bool g_scene_payload_persists = false;
thread_local int32_t g_scene_strip[2] = {0, 0};
void set_last_error(const std::string &msg);
}

namespace {
// Host array that lives in pinned memory when a CUDA driver is present (so UploadSceneD3D11 payloads
// are copied at PCIe speed without a staging pass) and in ordinary memory otherwise (CPU-only tests).
template <typename T>
struct HostBuffer {
    T *ptr = nullptr;
    size_t capacity = 0;
    bool pinned = false;
    HostBuffer() = default;
    HostBuffer(const HostBuffer &) = delete;
    HostBuffer &operator=(const HostBuffer &) = delete;
    ~HostBuffer() { release(); }
    void release() {
        if (!ptr) return;
        if (pinned)
            cudaFreeHost(ptr);
        else
            free(ptr);
        ptr = nullptr;
        capacity = 0;
    }
    void ensure(size_t n) {
        if (n <= capacity) return;
        release();
        size_t want = n + n / 8 + 64;
        void *p = nullptr;
        if (cudaMallocHost(&p, want * sizeof(T)) == cudaSuccess) {
            pinned = true;
        } else {
            (void)cudaGetLastError();
            p = malloc(want * sizeof(T));
            pinned = false;
        }
        ptr = static_cast<T *>(p);
        capacity = want;
    }
};
} // namespace

namespace {

struct RectF {
    float min_x, min_y, max_x, max_y;
};

inline float sse_min(float a, float b) { return a < b ? a : b; } // _mm_min_ps
inline float sse_max(float a, float b) { return a > b ? a : b; } // _mm_max_ps

inline RectF union_rect(RectF a, RectF b) { // geometry/src/rect.rs:114-120
    return RectF{sse_min(a.min_x, b.min_x), sse_min(a.min_y, b.min_y), sse_max(a.max_x, b.max_x),
                 sse_max(a.max_y, b.max_y)};
}

struct Transform {
    float m11 = 1, m21 = 0, m12 = 0, m22 = 1, tx = 0, ty = 0;
    bool is_identity() const { return m11 == 1 && m21 == 0 && m12 == 0 && m22 == 1 && tx == 0 && ty == 0; }
    // Transform2F * Vector2F (geometry/src/transform2d.rs:123-130,312-318)
    void apply(float x, float y, float &ox, float &oy) const {
        float hx = m11 * x, hy = m21 * x, hz = m12 * y, hw = m22 * y;
        ox = (hx + hz) + tx;
        oy = (hy + hw) + ty;
    }
    // Transform2F * RectF (geometry/src/transform2d.rs:329-338)
    RectF apply_rect(RectF r) const {
        float ulx, uly, urx, ury, llx, lly, lrx, lry;
        apply(r.min_x, r.min_y, ulx, uly);
        apply(r.max_x, r.min_y, urx, ury);
        apply(r.min_x, r.max_y, llx, lly);
        apply(r.max_x, r.max_y, lrx, lry);
        RectF o;
        o.min_x = sse_min(sse_min(sse_min(ulx, urx), llx), lrx);
        o.min_y = sse_min(sse_min(sse_min(uly, ury), lly), lry);
        o.max_x = sse_max(sse_max(sse_max(ulx, urx), llx), lrx);
        o.max_y = sse_max(sse_max(sse_max(uly, ury), lly), lry);
        return o;
    }
};

struct Path {
    uint32_t first_contour, end_contour;
    RectF bounds; // Outline::bounds(): min/max over all points of all non-empty contours
    uint16_t paint;
    uint8_t fill_rule, blend_mode;
    uint32_t clip_path;
    uint32_t segment_points;  // points SegmentsD3D11::add_path emits: all points + one closing copy per contour
    uint32_t segment_indices; // on-curve points = segments
};

// Scene::display_list (scene.rs:500-509): draw path ranges interleaved with render target pushes / pops.
struct DisplayItem {
    enum Kind : uint8_t { DRAW_PATHS, PUSH_RENDER_TARGET, POP_RENDER_TARGET } kind;
    uint32_t a, b; // DRAW_PATHS: draw path ids [a, b); PUSH_RENDER_TARGET: a = render target id
};
struct RenderTargetDesc { // scene.rs:493-497
    int32_t width, height;
};
// Paint::overlay (paint.rs:108-146): a pattern over a render target or an image (pattern.rs:52-140), or a gradient
// (content/src/gradient.rs).
struct PatternOverlay {
    enum Kind : uint8_t { RENDER_TARGET, IMAGE, GRADIENT } kind = RENDER_TARGET;
    uint32_t render_target = 0;
    Transform transform; // pattern space (render target / image pixels) -> scene space; radial gradients: Radial.transform
    PFFilter filter;
    // IMAGE
    std::vector<PFColorU> pixels;
    int32_t width = 0, height = 0;
    uint32_t pattern_flags = 0; // PF_PATTERN_FLAG_*
    // GRADIENT
    uint32_t gradient_kind = 0, gradient_wrap = 0;
    float from[2] = {0, 0}, to[2] = {0, 0}, radii[2] = {0, 0};
    std::vector<PFColorStop> stops;
    // assigned by the build (Palette::assign_paint_locations, paint.rs:456-595)
    uint32_t page = 0;
    int32_t row = 0;          // gradients: the row of the 256 x 256 gradient tile
    int32_t page_w = 0, page_h = 0;
};
// One DrawTilesD3D11 batch of a scene with a display list (the general builder below).
struct GeneralBatch {
    std::vector<PFPropagateMetadataD3D11> propagate_metadata;
    std::vector<PFDiceMetadataD3D11> dice_metadata;
    std::vector<PFTilePathInfoD3D11> tile_path_info;
    uint32_t tile_count = 0, column_count = 0, segment_count = 0;
    uint32_t clipped_paths = 0, clipped_tiles = 0; // paths with a clip path in this batch, and their tiles
    bool has_color_texture = false;
    PFTileBatchTexture color_texture{0, 0, 0};
};

// Gradient::sample (content/src/gradient.rs:188-211): the colour at t in [0, 1], stops sorted by offset.
PFColorU gradient_sample(const std::vector<PFColorStop> &stops, float t) {
    if (stops.empty()) return PFColorU{0, 0, 0, 0};
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    const size_t last = stops.size() - 1;
    // binary_search_by(|stop| if stop.offset < t || stop.offset == 0.0 { Less } else { Greater }): the first stop
    // that is neither before t nor at offset zero
    size_t lo = 0, hi = stops.size();
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        if (stops[mid].offset < t || stops[mid].offset == 0.0f) lo = mid + 1;
        else hi = mid;
    }
    const size_t upper = lo < last ? lo : last;
    const size_t lower = upper > 0 ? upper - 1 : upper;
    const PFColorStop &a = stops[lower], &b = stops[upper];
    const float denom = b.offset - a.offset;
    if (denom == 0.0f) return a.color;
    float ratio = (t - a.offset) / denom;
    if (ratio > 1.0f) ratio = 1.0f;
    // ColorU::to_f32().lerp(..).to_u8(): (x / 255 + (y / 255 - x / 255) * ratio) * 255 through F32x4::to_i32x4, i.e.
    // cvtps: round to nearest even (color/src/lib.rs:70-73,163-171)
    auto lerp8 = [&](uint8_t x, uint8_t y) {
        const float fx = (float)x * (1.0f / 255.0f), fy = (float)y * (1.0f / 255.0f);
        const float v = (fx + (fy - fx) * ratio) * 255.0f;
        const float r = nearbyintf(v);
        return (uint8_t)(r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r));
    };
    return PFColorU{lerp8(a.color.r, b.color.r), lerp8(a.color.g, b.color.g), lerp8(a.color.b, b.color.b), lerp8(a.color.a, b.color.a)};
}

std::atomic<uint32_t> g_next_scene_id{0}; // NEXT_SCENE_ID, scene.rs:34

} // namespace

struct PFScene {
    std::vector<PFVector2F> points;
    std::vector<uint8_t> flags;
    std::vector<uint32_t> contour_offsets{0};
    std::vector<Path> draw_paths, clip_paths;
    std::vector<RectF> draw_bounds; // every draw path's bounds again, packed: what a strip build scans (16 bytes a path)
    std::vector<PFColorU> paints;
    std::unordered_map<uint32_t, uint16_t> paint_cache; // Palette::push_paint dedup (paint.rs:115-131)
    // Render targets, pattern paints over them (paint id -> overlay) and the display list. A scene without
    // render targets has the display list [DrawPaths(0..n)] and takes the cached single-batch build.
    std::vector<RenderTargetDesc> render_targets;
    std::unordered_map<uint16_t, PatternOverlay> overlays;
    std::vector<DisplayItem> display_list;
    std::vector<GeneralBatch> general_batches; // scratch of the general builder (kept alive while commands are sent)
    bool any_blend = false;                    // some draw path blends with something other than SrcOver
    std::unordered_map<uint32_t, uint16_t> blend_entries; // paint | blend mode << 16 -> texture metadata entry
    std::vector<uint8_t> strip_included;       // per draw path: has a tile in the renderer's strip
    std::vector<uint32_t> strip_ids;           // ... and those paths' ids, ascending (what a strip build iterates over)
    std::vector<std::vector<PFColorU>> gradient_tiles; // scratch: texels of the gradient pages of this build
    RectF bounds{0, 0, 0, 0};
    RectF view_box{0, 0, 0, 0};
    uint32_t id;
    uint32_t epoch = 0;

    // Set while a renderer may still be copying seg_points / seg_indices (payload_persists uploads).
    cudaEvent_t borrowed_event = nullptr;
    int borrowed_device = -1;
    bool borrowed = false;

    // Scratch reused across builds.
    HostBuffer<PFVector2F> seg_points;
    HostBuffer<PFSegmentIndicesD3D11> seg_indices;
    size_t seg_point_count = 0, seg_index_count = 0;
    // The same for the clip paths (BuiltSegments::from_scene builds both, builder.rs:801-815).
    HostBuffer<PFVector2F> clip_seg_points;
    HostBuffer<PFSegmentIndicesD3D11> clip_seg_indices;
    size_t clip_seg_point_count = 0, clip_seg_index_count = 0;
    std::vector<uint32_t> clip_seg_path_offsets, clip_segment_ranges;
    // Clip batch arrays (PrepareClipTilesD3D11) and the clip path -> batch index map of the last build.
    std::vector<PFPropagateMetadataD3D11> clip_propagate_metadata;
    std::vector<PFDiceMetadataD3D11> clip_dice_metadata;
    std::vector<PFTilePathInfoD3D11> clip_tile_path_info;
    std::vector<uint32_t> clip_batch_index;
    uint32_t built_clip_tile_count = 0, built_clip_segment_count = 0, built_clipped_path_count = 0,
             built_clipped_tile_count = 0;
    std::vector<uint32_t> seg_path_offsets; // per-path output offsets of build_segments
    // BuildOptions with a dilation (prepare_paths): the points after Scene::apply_render_options, same layout as
    // `points`, and the bounds the CPU tiler sees for every path. segments_key identifies what the segment arrays
    // currently hold (scene epoch x prepared options); upload_serial counts their rewrites for the sinks.
    std::vector<PFVector2F> prepared_points;
    std::vector<RectF> prepared_draw_bounds, prepared_clip_bounds;
    uint64_t segments_key = 0;
    uint32_t upload_serial = 0;
    std::vector<PFRectI> path_tile_rects;   // scratch of the batch build (kept == rect non-empty)
    std::vector<uint32_t> path_batch_offsets;
    std::vector<uint32_t> draw_segment_ranges; // [n_draw][2]
    std::vector<PFPropagateMetadataD3D11> propagate_metadata;
    std::vector<PFDiceMetadataD3D11> dice_metadata;
    std::vector<PFTilePathInfoD3D11> tile_path_info;
    std::vector<PFTextureMetadataEntry> texture_metadata;
    // Key of the build whose batch arrays are still in the vectors above (0 = none).
    uint64_t built_key = 0;
    uint64_t built_paint_key = 0;
    uint32_t built_tile_count = 0, built_segment_count = 0;

    PFScene() : id(g_next_scene_id.fetch_add(1)) {}
};

struct PFRenderTransform {
    Transform t;
};
struct PFBuildOptions {
    Transform transform; // RenderTransform::Transform2D; default identity (options.rs:80-85)
    float dilation[2] = {0, 0};
    bool subpixel_aa_enabled = false;
};

namespace {
// Scene::push_draw_path (scene.rs:87-92): extend the trailing DrawPaths item or start a new one.
void note_draw_paths_pushed(PFScene *s, uint32_t count) {
    const uint32_t end = (uint32_t)s->draw_paths.size();
    if (!s->display_list.empty() && s->display_list.back().kind == DisplayItem::DRAW_PATHS)
        s->display_list.back().b = end;
    else
        s->display_list.push_back(DisplayItem{DisplayItem::DRAW_PATHS, end - count, end});
}

// Transform2F::inverse (geometry/src/transform2d.rs:289-293) and a * b (apply b first), in f32.
Transform transform_inverse(const Transform &t) {
    const float det = t.m11 * t.m22 - t.m12 * t.m21;
    const float inv = 1.0f / det;
    Transform r;
    r.m11 = t.m22 * inv, r.m12 = -t.m12 * inv, r.m21 = -t.m21 * inv, r.m22 = t.m11 * inv;
    r.tx = -(r.m11 * t.tx + r.m12 * t.ty);
    r.ty = -(r.m21 * t.tx + r.m22 * t.ty);
    return r;
}
Transform transform_mul(const Transform &a, const Transform &b) {
    Transform r;
    r.m11 = a.m11 * b.m11 + a.m12 * b.m21, r.m12 = a.m11 * b.m12 + a.m12 * b.m22;
    r.m21 = a.m21 * b.m11 + a.m22 * b.m21, r.m22 = a.m21 * b.m12 + a.m22 * b.m22;
    r.tx = a.m11 * b.tx + a.m12 * b.ty + a.tx;
    r.ty = a.m21 * b.tx + a.m22 * b.ty + a.ty;
    return r;
}
} // namespace

namespace pf {
void scene_note_borrowed(PFScene *s, cudaStream_t stream, int device) {
    if (s->borrowed_event && s->borrowed_device != device) {
        // the scene moved to a renderer on another GPU: events belong to the device they were created on
        if (s->borrowed) cudaEventSynchronize(s->borrowed_event);
        cudaEventDestroy(s->borrowed_event);
        s->borrowed_event = nullptr;
    }
    if (!s->borrowed_event) {
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&s->borrowed_event, cudaEventDisableTiming));
        s->borrowed_device = device;
    } else if (s->borrowed) {
        // A second renderer borrows the same arrays while an earlier copy may still be in flight on another
        // stream (the arrays are not rebuilt for a new sink when the scene has not changed): one event can only
        // stand for one of them, so finish the earlier one first. Never taken with one renderer per scene.
        if (cudaEventSynchronize(s->borrowed_event) != cudaSuccess) (void)cudaGetLastError();
    }
    PF_CUDA_CHECK(cudaEventRecord(s->borrowed_event, stream));
    s->borrowed = true;
}
} // namespace pf

namespace {

// Appends an outline's contours to the scene pools and returns its bounds
// (Contour::push_point with update_bounds, content/src/outline.rs:560-573; Outline::push_contour
// :180-192 drops empty contours).
Path append_outline(PFScene *s, const PFVector2F *points, const uint8_t *point_flags,
                    const uint32_t *contour_offsets, uint32_t contour_count) {
    Path p{};
    p.first_contour = (uint32_t)s->contour_offsets.size() - 1;
    bool have_bounds = false;
    RectF bounds{0, 0, 0, 0};
    for (uint32_t c = 0; c < contour_count; c++) {
        uint32_t p0 = contour_offsets[c], p1 = contour_offsets[c + 1];
        if (p0 == p1) continue;
        RectF cb{points[p0].x, points[p0].y, points[p0].x, points[p0].y};
        p.segment_points += p1 - p0 + 1;
        for (uint32_t i = p0; i < p1; i++) {
            s->points.push_back(points[i]);
            s->flags.push_back(point_flags[i]);
            p.segment_indices += (point_flags[i] & (PF_POINT_FLAGS_CONTROL_POINT_0 | PF_POINT_FLAGS_CONTROL_POINT_1)) ? 0u : 1u;
            cb.min_x = sse_min(cb.min_x, points[i].x);
            cb.min_y = sse_min(cb.min_y, points[i].y);
            cb.max_x = sse_max(cb.max_x, points[i].x);
            cb.max_y = sse_max(cb.max_y, points[i].y);
        }
        s->contour_offsets.push_back((uint32_t)s->points.size());
        bounds = have_bounds ? union_rect(bounds, cb) : cb;
        have_bounds = true;
    }
    p.end_contour = (uint32_t)s->contour_offsets.size() - 1;
    p.bounds = bounds;
    return p;
}

// Arguments of the push calls that would otherwise be read out of bounds later (the reference panics on the
// same mistakes, e.g. an unknown PaintId indexes Palette::paints in build_paint_info).
bool outline_args_ok(const PFScene *s, const PFVector2F *points, const uint8_t *point_flags,
                     const uint32_t *contour_offsets, uint32_t contour_count, const char *who) {
    if (!s || (contour_count && (!contour_offsets || !points || !point_flags))) {
        pf::set_last_error(std::string(who) + ": null argument");
        return false;
    }
    for (uint32_t c = 0; c < contour_count; c++) {
        if (contour_offsets[c] > contour_offsets[c + 1]) {
            pf::set_last_error(std::string(who) + ": contour_offsets must not decrease");
            return false;
        }
    }
    return true;
}

using pf::parallel_ranges;

// BuiltSegments::from_scene + SegmentsD3D11::add_path (renderer/src/builder.rs:777-841) for the draw
// paths: every contour's points followed by its first point again (implicit close), one index entry
// per on-curve point, flagged quadratic / cubic by the control points that follow it. Output
// positions are prefix sums of per-path counts, so paths are filled in parallel.
// The previous upload of the segment arrays may still be in flight on a renderer's stream.
void wait_for_borrowers(PFScene *s) {
    if (!s->borrowed) return;
    s->borrowed = false;
    if (cudaEventSynchronize(s->borrowed_event) != cudaSuccess) (void)cudaGetLastError();
}

// SegmentsD3D11::add_path for every path of one kind (builder.rs:817-857).
void build_path_segments(PFScene *s, const PFVector2F *scene_points, const std::vector<Path> &paths,
                         HostBuffer<PFVector2F> &seg_points,
                         HostBuffer<PFSegmentIndicesD3D11> &seg_indices, size_t &seg_point_count, size_t &seg_index_count,
                         std::vector<uint32_t> &path_offsets, std::vector<uint32_t> &segment_ranges,
                         const std::vector<uint32_t> *ids = nullptr) {
    // `ids` (optional, ascending): the paths to emit. The others are not touched at all — a renderer that owns a strip
    // of the frame only needs the paths that reach it, and at N ranks even reading 100k path records per frame on a
    // rank's share of the host cores would cost more than the strip's own work.
    const size_t n_paths = paths.size();
    const size_t n_items = ids ? ids->size() : n_paths;
    const uint8_t ctrl_mask = PF_POINT_FLAGS_CONTROL_POINT_0 | PF_POINT_FLAGS_CONTROL_POINT_1;
    segment_ranges.resize(2 * n_paths);
    // One path: its contours' points (each contour closed by its first point again, builder.rs:835) from `wp`, one index
    // entry per on-curve point from `wi`; both cursors are left at the end of the path.
    auto emit_path = [&](size_t pi, PFVector2F *out_points, PFSegmentIndicesD3D11 *out_indices, size_t &wp, size_t &wi) {
        const Path &path = paths[pi];
        segment_ranges[2 * pi] = (uint32_t)wi;
        for (uint32_t c = path.first_contour; c < path.end_contour; c++) {
            const uint32_t p0 = s->contour_offsets[c], point_count = s->contour_offsets[c + 1] - p0;
            const uint8_t *flags = s->flags.data() + p0;
            const PFVector2F *pts = scene_points + p0;
            memcpy(out_points + wp, pts, (size_t)point_count * sizeof(PFVector2F));
            for (uint32_t i = 0; i < point_count; i++) {
                if (flags[i] & ctrl_mask) continue;
                uint32_t f = 0;
                if (i + 1 < point_count && (flags[i + 1] & PF_POINT_FLAGS_CONTROL_POINT_0))
                    f = (i + 2 < point_count && (flags[i + 2] & PF_POINT_FLAGS_CONTROL_POINT_1))
                            ? PF_CURVE_IS_CUBIC
                            : PF_CURVE_IS_QUADRATIC;
                out_indices[wi++] = PFSegmentIndicesD3D11{(uint32_t)(wp + i), f};
            }
            wp += point_count;
            out_points[wp++] = pts[0]; // implicit close: the first point again (builder.rs:835)
        }
        segment_ranges[2 * pi + 1] = (uint32_t)wi;
    };
    uint32_t np = 0, ni = 0;
    if (ids) {
        // A strip's paths lie far apart in the scene's arrays, so every visit of a path record is a cache miss unless
        // it is fetched ahead. Two-level scan over fixed chunks of the id list, both levels on the worker pool: pass A
        // sums what a chunk's paths add to the two output cursors, pass B starts from the prefix over chunks and
        // advances the cursors path by path (a chunk's paths are contiguous in the output) — no per-path offset
        // arrays, no serial pass over the records, and a chunk's records are still in its thread's cache in pass B.
        struct ChunkSum {
            uint32_t points = 0, indices = 0;
            uint32_t pad[14]; // one cache line per chunk
        };
        const size_t chunks = pf::chunk_count(n_items, 1024);
        std::vector<ChunkSum> sums(chunks + 1);
        const uint32_t *id = ids->data();
        pf::parallel_chunks(n_items, chunks, [&](size_t c, size_t begin, size_t end) {
            uint32_t points = 0, indices = 0;
            for (size_t k = begin; k < end; k++) {
                if (k + 16 < end) __builtin_prefetch(&paths[id[k + 16]]);
                points += paths[id[k]].segment_points;
                indices += paths[id[k]].segment_indices;
            }
            sums[c + 1].points = points, sums[c + 1].indices = indices;
        });
        for (size_t c = 1; c <= chunks; c++) // exclusive prefix: sums[c] = totals of the chunks before c
            sums[c].points += sums[c - 1].points, sums[c].indices += sums[c - 1].indices;
        np = sums[chunks].points, ni = sums[chunks].indices;
        seg_points.ensure((size_t)np + 1);
        seg_indices.ensure((size_t)ni + 1);
        PFVector2F *out_points = seg_points.ptr;
        PFSegmentIndicesD3D11 *out_indices = seg_indices.ptr;
        pf::parallel_chunks(n_items, chunks, [&](size_t c, size_t begin, size_t end) {
            size_t wp = sums[c].points, wi = sums[c].indices;
            for (size_t k = begin; k < end; k++) {
                // The chain path record -> contour offsets -> points + flags is fetched a few paths ahead, one link
                // per step.
                if (k + 12 < end) __builtin_prefetch(&paths[id[k + 12]]);
                if (k + 8 < end) __builtin_prefetch(&s->contour_offsets[paths[id[k + 8]].first_contour]);
                if (k + 4 < end) {
                    const uint32_t p_ahead = s->contour_offsets[paths[id[k + 4]].first_contour];
                    __builtin_prefetch(scene_points + p_ahead);
                    __builtin_prefetch(scene_points + p_ahead + 8);
                    __builtin_prefetch(scene_points + p_ahead + 16);
                    __builtin_prefetch(s->flags.data() + p_ahead);
                }
                emit_path(id[k], out_points, out_indices, wp, wi);
            }
        });
    } else {
        // Every path: output positions are prefix sums of the per-path counts (a sequential pass over the records),
        // so that the paths are then filled in parallel.
        path_offsets.resize(2 * (n_paths + 1));
        uint32_t *point_off = path_offsets.data(), *index_off = point_off + (n_paths + 1);
        for (size_t pi = 0; pi < n_paths; pi++) {
            point_off[pi] = np, index_off[pi] = ni;
            np += paths[pi].segment_points;
            ni += paths[pi].segment_indices;
        }
        point_off[n_paths] = np, index_off[n_paths] = ni;
        seg_points.ensure((size_t)np + 1);
        seg_indices.ensure((size_t)ni + 1);
        PFVector2F *out_points = seg_points.ptr;
        PFSegmentIndicesD3D11 *out_indices = seg_indices.ptr;
        parallel_ranges(n_paths, 4096, [&](size_t begin, size_t end) {
            for (size_t pi = begin; pi < end; pi++) {
                size_t wp = point_off[pi], wi = index_off[pi];
                emit_path(pi, out_points, out_indices, wp, wi);
            }
        });
    }
    seg_point_count = np;
    seg_index_count = ni;
}

// `scene_points`: the scene's own points, or their prepared copy.
void build_segments(PFScene *s, const PFVector2F *scene_points, const std::vector<uint32_t> *draw_ids = nullptr) {
    wait_for_borrowers(s);
    build_path_segments(s, scene_points, s->draw_paths, s->seg_points, s->seg_indices, s->seg_point_count, s->seg_index_count,
                        s->seg_path_offsets, s->draw_segment_ranges, draw_ids);
    build_path_segments(s, scene_points, s->clip_paths, s->clip_seg_points, s->clip_seg_indices, s->clip_seg_point_count,
                        s->clip_seg_index_count, s->clip_seg_path_offsets, s->clip_segment_ranges);
}

// Scene::apply_render_options, 2-D branch (renderer/src/scene.rs:249-270), for every path of one kind, when the
// build options carry a dilation: transform the points (Outline::transform, outline.rs:208-221, skipped for the
// identity), recompute the bounds over all points, dilate (dilate.cpp) and grow the bounds by the amount
// (Outline::dilate, outline.rs:243-249). The reference's D3D11 builder drops the dilation (options.rs:165-180 hands
// only the transform to the GPU); the CPU tiler, which is the parity target, applies it here, so the host does the
// same and the device then dices already-prepared points under an identity transform.
void prepare_paths(PFScene *s, const std::vector<Path> &paths, const Transform &xf, const float dilation[2],
                   std::vector<RectF> &bounds_out) {
    bounds_out.resize(paths.size());
    const bool identity = xf.is_identity();
    PFVector2F *out = s->prepared_points.data();
    parallel_ranges(paths.size(), 1024, [&](size_t begin, size_t end) {
        for (size_t pi = begin; pi < end; pi++) {
            const Path &path = paths[pi];
            const uint32_t *offsets = s->contour_offsets.data() + path.first_contour;
            const uint32_t contours = path.end_contour - path.first_contour;
            RectF bounds{0, 0, 0, 0};
            for (uint32_t c = 0; c < contours; c++) {
                RectF cb{0, 0, 0, 0};
                for (uint32_t i = offsets[c]; i < offsets[c + 1]; i++) {
                    PFVector2F p = s->points[i];
                    if (!identity) xf.apply(p.x, p.y, p.x, p.y);
                    out[i] = p;
                    if (i == offsets[c]) cb = RectF{p.x, p.y, p.x, p.y};
                    cb = RectF{sse_min(cb.min_x, p.x), sse_min(cb.min_y, p.y), sse_max(cb.max_x, p.x), sse_max(cb.max_y, p.y)};
                }
                bounds = c == 0 ? cb : union_rect(bounds, cb); // the scene holds no empty contours (append_outline)
            }
            pf::dilate_outline(out, offsets, contours, dilation[0], dilation[1]);
            bounds_out[pi] = RectF{bounds.min_x - dilation[0], bounds.min_y - dilation[1], bounds.max_x + dilation[0],
                                   bounds.max_y + dilation[1]}; // RectF::dilate (rect.rs:171-180)
        }
    });
}

// RectF::intersection (geometry/src/rect.rs:122-137), strict comparisons.
bool rect_intersection(RectF a, RectF b, RectF &out) {
    if (!(a.min_x < b.max_x && a.min_y < b.max_y && b.min_x < a.max_x && b.min_y < a.max_y)) return false;
    out = RectF{sse_max(a.min_x, b.min_x), sse_max(a.min_y, b.min_y), sse_min(a.max_x, b.max_x),
                sse_min(a.max_y, b.max_y)};
    return true;
}

// splitmix64-style mixing for content keys.
uint64_t mix_key(uint64_t h, uint64_t v) {
    uint64_t z = h ^ (v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

PFRenderCommand make_command(uint32_t kind) {
    PFRenderCommand c;
    memset(&c, 0, sizeof(c));
    c.kind = kind;
    return c;
}

// The clip batch of a build (PrepareClipTilesD3D11): fills s->clip_batch_index / clip_propagate_metadata /
// clip_dice_metadata / clip_tile_path_info and the batch totals.
PFCudaStatus build_clip_batch(PFScene *s, bool prepared, const Transform &xf, const RectF &view_box) {
    // Clip batch (add_clip_path_to_batch, builder.rs:1124-1175): the clip paths some draw path uses, in
    // order of first use, one level deep. The tile rect follows the CPU tiler, which is the parity
    // target: outline bounds ∩ view box (Tiler::new, tiler.rs:47-50), an empty rect when they miss.
    s->clip_batch_index.assign(s->clip_paths.size(), PF_PATH_INDEX_NONE);
    s->clip_propagate_metadata.clear();
    s->clip_dice_metadata.clear();
    s->clip_tile_path_info.clear();
    uint32_t clip_tiles = 0, clip_columns = 0, clip_segments = 0;
    for (const Path &p : s->draw_paths) {
        if (p.clip_path == PF_CLIP_PATH_NONE) continue;
        if (p.clip_path >= s->clip_paths.size()) {
            pf::set_last_error("draw path refers to a clip path that does not exist");
            return PF_CUDA_ERROR_INVALID_ARGUMENT;
        }
        if (s->clip_batch_index[p.clip_path] != PF_PATH_INDEX_NONE) continue;
        const Path &cp = s->clip_paths[p.clip_path];
        if (cp.clip_path != PF_CLIP_PATH_NONE) {
            pf::set_last_error("nested clip paths are not implemented");
            return PF_CUDA_ERROR_UNSUPPORTED;
        }
        PFRectI tile_rect{{0, 0}, {0, 0}};
        RectF bounds = prepared ? s->prepared_clip_bounds[p.clip_path]
                                : xf.is_identity() ? cp.bounds : xf.apply_rect(cp.bounds);
        RectF clipped;
        if (cp.first_contour != cp.end_contour && rect_intersection(bounds, view_box, clipped)) {
            const float k = 1.0f / 16.0f;
            tile_rect.origin.x = (int32_t)floorf(clipped.min_x * k);
            tile_rect.origin.y = (int32_t)floorf(clipped.min_y * k);
            tile_rect.lower_right.x = (int32_t)ceilf(clipped.max_x * k);
            tile_rect.lower_right.y = (int32_t)ceilf(clipped.max_y * k);
        }
        const uint32_t bi = (uint32_t)s->clip_propagate_metadata.size();
        s->clip_batch_index[p.clip_path] = bi;
        const uint32_t w = (uint32_t)(tile_rect.lower_right.x - tile_rect.origin.x),
                       h = (uint32_t)(tile_rect.lower_right.y - tile_rect.origin.y);
        PFPropagateMetadataD3D11 pm;
        memset(&pm, 0, sizeof(pm));
        pm.tile_rect = tile_rect;
        pm.tile_offset = clip_tiles;
        pm.path_index = bi;
        pm.z_write = 0;
        pm.clip_path_index = PF_PATH_INDEX_NONE;
        pm.backdrop_offset = clip_columns;
        s->clip_propagate_metadata.push_back(pm);
        s->clip_dice_metadata.push_back(
            PFDiceMetadataD3D11{p.clip_path, s->clip_segment_ranges[2 * p.clip_path], clip_segments, 0});
        PFTilePathInfoD3D11 tp;
        tp.tile_min_x = (int16_t)tile_rect.origin.x;
        tp.tile_min_y = (int16_t)tile_rect.origin.y;
        tp.tile_max_x = (int16_t)tile_rect.lower_right.x;
        tp.tile_max_y = (int16_t)tile_rect.lower_right.y;
        tp.first_tile_index = clip_tiles;
        tp.color = 0; // TilingPathInfo::Clip: paint 0, ctrl 0 (builder.rs:423-428, tiles.rs:45-61)
        tp.ctrl = 0;
        tp.backdrop = 0;
        s->clip_tile_path_info.push_back(tp);
        clip_tiles += w * h;
        clip_columns += w;
        clip_segments += s->clip_segment_ranges[2 * p.clip_path + 1] - s->clip_segment_ranges[2 * p.clip_path];
    }
    s->built_clip_tile_count = clip_tiles;
    s->built_clip_segment_count = clip_segments;
    return PF_CUDA_OK;
}

// Sends the clip batch built by build_clip_batch (clip batches are prepared before the draw batches that use them,
// builder.rs:1098-1104).
PFCudaStatus send_clip_batch(PFScene *s, const Transform &xf, PFRenderCommandListenerFn listener, void *userdata) {
    PFRenderCommand prepare = make_command(PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11);
    PFTileBatchDataD3D11 &b = prepare.u.prepare_clip_tiles_d3d11.batch;
    b.batch_id = 0; // clip level 0
    b.path_count = (uint32_t)s->clip_propagate_metadata.size();
    b.tile_count = s->built_clip_tile_count;
    b.segment_count = s->built_clip_segment_count;
    b.prepare_info.backdrops = nullptr;
    b.prepare_info.backdrop_count = 0;
    b.prepare_info.propagate_metadata = s->clip_propagate_metadata.data();
    b.prepare_info.dice_metadata = s->clip_dice_metadata.data();
    b.prepare_info.tile_path_info = s->clip_tile_path_info.data();
    b.prepare_info.transform.matrix = PFMatrix2x2F{xf.m11, xf.m12, xf.m21, xf.m22};
    b.prepare_info.transform.vector = PFVector2F{xf.tx, xf.ty};
    b.path_source = PF_PATH_SOURCE_CLIP;
    b.has_clipped_path_info = 0;
    b.content_key = 0;
    return listener(&prepare, userdata);
}


} // namespace

extern "C" {

PFSceneRef PFSceneCreate(void) { return new PFScene(); }
void PFSceneDestroy(PFSceneRef scene) {
    if (!scene) return;
    wait_for_borrowers(scene);
    if (scene->borrowed_event) cudaEventDestroy(scene->borrowed_event);
    delete scene;
}

// Scene: Clone (scene.rs:37): the content — outlines, paths, palette, display list, bounds, view box — with the same
// id and epoch; the build scratch (segment arrays, batch arrays) starts empty in the copy.
PFSceneRef PFSceneClone(PFSceneRef s) {
    if (!s) return nullptr;
    PFScene *c = new PFScene();
    c->points = s->points, c->flags = s->flags, c->contour_offsets = s->contour_offsets;
    c->draw_paths = s->draw_paths, c->clip_paths = s->clip_paths, c->draw_bounds = s->draw_bounds;
    c->paints = s->paints, c->paint_cache = s->paint_cache;
    c->render_targets = s->render_targets, c->overlays = s->overlays, c->display_list = s->display_list;
    c->any_blend = s->any_blend;
    c->bounds = s->bounds, c->view_box = s->view_box;
    c->id = s->id, c->epoch = s->epoch;
    return c;
}

void PFSceneSetViewBox(PFSceneRef s, const PFRectF *vb) {
    s->view_box = RectF{vb->origin.x, vb->origin.y, vb->lower_right.x, vb->lower_right.y};
    s->epoch++;
}
void PFSceneGetViewBox(PFSceneRef s, PFRectF *vb) {
    vb->origin = PFVector2F{s->view_box.min_x, s->view_box.min_y};
    vb->lower_right = PFVector2F{s->view_box.max_x, s->view_box.max_y};
}
void PFSceneGetBounds(PFSceneRef s, PFRectF *b) {
    b->origin = PFVector2F{s->bounds.min_x, s->bounds.min_y};
    b->lower_right = PFVector2F{s->bounds.max_x, s->bounds.max_y};
}

uint16_t PFScenePushPaint(PFSceneRef s, const PFColorU *color) {
    uint32_t key = (uint32_t)color->r | ((uint32_t)color->g << 8) | ((uint32_t)color->b << 16) | ((uint32_t)color->a << 24);
    auto it = s->paint_cache.find(key);
    if (it != s->paint_cache.end()) return it->second;
    uint16_t id = (uint16_t)s->paints.size(); // PaintId(u16), paint.rs:81
    s->paints.push_back(*color);
    s->paint_cache.emplace(key, id);
    s->epoch++;
    return id;
}

uint32_t PFScenePushDrawPath(PFSceneRef s, const PFVector2F *points, const uint8_t *point_flags,
                             const uint32_t *contour_offsets, uint32_t contour_count, uint16_t paint_id,
                             uint8_t fill_rule, uint8_t blend_mode, uint32_t clip_path_id) {
    if (!outline_args_ok(s, points, point_flags, contour_offsets, contour_count, "PFScenePushDrawPath")) return PF_PATH_INDEX_NONE;
    if (paint_id >= s->paints.size()) {
        pf::set_last_error("PFScenePushDrawPath: unknown paint id");
        return PF_PATH_INDEX_NONE;
    }
    if (fill_rule != PF_FILL_RULE_WINDING && fill_rule != PF_FILL_RULE_EVEN_ODD) {
        pf::set_last_error("PFScenePushDrawPath: unknown fill rule");
        return PF_PATH_INDEX_NONE;
    }
    Path p = append_outline(s, points, point_flags, contour_offsets, contour_count);
    p.paint = paint_id;
    p.fill_rule = fill_rule;
    p.blend_mode = blend_mode;
    if (blend_mode != PF_BLEND_MODE_SRC_OVER) s->any_blend = true;
    p.clip_path = clip_path_id;
    s->bounds = union_rect(s->bounds, p.bounds); // scene.rs:84-86
    s->draw_paths.push_back(p);
    s->draw_bounds.push_back(p.bounds);
    note_draw_paths_pushed(s, 1);
    s->epoch++;
    return (uint32_t)s->draw_paths.size() - 1;
}

uint32_t PFScenePushClipPath(PFSceneRef s, const PFVector2F *points, const uint8_t *point_flags,
                             const uint32_t *contour_offsets, uint32_t contour_count, uint8_t fill_rule,
                             uint32_t clip_path_id) {
    if (!outline_args_ok(s, points, point_flags, contour_offsets, contour_count, "PFScenePushClipPath")) return PF_PATH_INDEX_NONE;
    if (fill_rule != PF_FILL_RULE_WINDING && fill_rule != PF_FILL_RULE_EVEN_ODD) {
        pf::set_last_error("PFScenePushClipPath: unknown fill rule");
        return PF_PATH_INDEX_NONE;
    }
    Path p = append_outline(s, points, point_flags, contour_offsets, contour_count);
    p.paint = 0;
    p.fill_rule = fill_rule;
    p.blend_mode = PF_BLEND_MODE_SRC_OVER;
    p.clip_path = clip_path_id;
    s->bounds = union_rect(s->bounds, p.bounds);
    s->clip_paths.push_back(p);
    s->epoch++;
    return (uint32_t)s->clip_paths.size() - 1;
}

PFCudaStatus PFScenePushDrawPaths(PFSceneRef s, const PFVector2F *points, const uint8_t *point_flags,
                                  size_t point_count, const uint32_t *contour_offsets, size_t contour_count,
                                  const uint32_t *path_contour_offsets, size_t path_count,
                                  const uint16_t *paint_ids, const uint8_t *fill_rules,
                                  const uint32_t *clip_path_ids) {
    // Validate everything before touching the scene, so that a refused call leaves it unchanged.
    if (!s || (path_count && (!path_contour_offsets || !paint_ids || !fill_rules)) ||
        (contour_count && (!contour_offsets || !points || !point_flags))) {
        pf::set_last_error("PFScenePushDrawPaths: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    if (contour_count && contour_offsets[contour_count] != point_count) {
        pf::set_last_error("PFScenePushDrawPaths: contour_offsets do not cover the points");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    for (size_t c = 0; c < contour_count; c++) {
        if (contour_offsets[c] > contour_offsets[c + 1]) {
            pf::set_last_error("PFScenePushDrawPaths: contour_offsets must not decrease");
            return PF_CUDA_ERROR_INVALID_ARGUMENT;
        }
    }
    for (size_t i = 0; i < path_count; i++) {
        if (path_contour_offsets[i] > path_contour_offsets[i + 1] || path_contour_offsets[i + 1] > contour_count) {
            pf::set_last_error("PFScenePushDrawPaths: path_contour_offsets do not index the contours");
            return PF_CUDA_ERROR_INVALID_ARGUMENT;
        }
        if (paint_ids[i] >= s->paints.size()) {
            pf::set_last_error("PFScenePushDrawPaths: unknown paint id");
            return PF_CUDA_ERROR_INVALID_ARGUMENT;
        }
        if (fill_rules[i] != PF_FILL_RULE_WINDING && fill_rules[i] != PF_FILL_RULE_EVEN_ODD) {
            pf::set_last_error("PFScenePushDrawPaths: unknown fill rule");
            return PF_CUDA_ERROR_INVALID_ARGUMENT;
        }
    }
    s->points.reserve(s->points.size() + point_count);
    s->flags.reserve(s->flags.size() + point_count);
    s->contour_offsets.reserve(s->contour_offsets.size() + contour_count);
    s->draw_paths.reserve(s->draw_paths.size() + path_count);
    for (size_t i = 0; i < path_count; i++) {
        uint32_t c0 = path_contour_offsets[i], c1 = path_contour_offsets[i + 1];
        Path p = append_outline(s, points, point_flags, contour_offsets + c0, c1 - c0);
        p.paint = paint_ids[i];
        p.fill_rule = fill_rules[i];
        p.blend_mode = PF_BLEND_MODE_SRC_OVER;
        p.clip_path = clip_path_ids ? clip_path_ids[i] : PF_CLIP_PATH_NONE;
        s->bounds = union_rect(s->bounds, p.bounds);
        s->draw_paths.push_back(p);
        s->draw_bounds.push_back(p.bounds);
    }
    if (path_count) note_draw_paths_pushed(s, (uint32_t)path_count);
    s->epoch++;
    return PF_CUDA_OK;
}

uint32_t PFScenePushRenderTarget(PFSceneRef s, int32_t width, int32_t height) {
    if (!s || width <= 0 || height <= 0 || (int64_t)width * height > (1ll << 30)) {
        pf::set_last_error("PFScenePushRenderTarget: bad size");
        return PF_PATH_INDEX_NONE;
    }
    const uint32_t id = (uint32_t)s->render_targets.size(); // Palette::push_render_target (paint.rs:133-137)
    s->render_targets.push_back(RenderTargetDesc{width, height});
    s->display_list.push_back(DisplayItem{DisplayItem::PUSH_RENDER_TARGET, id, 0});
    s->epoch++;
    return id;
}

void PFScenePopRenderTarget(PFSceneRef s) {
    s->display_list.push_back(DisplayItem{DisplayItem::POP_RENDER_TARGET, 0, 0});
    s->epoch++;
}

uint16_t PFScenePushPaintRenderTargetPattern(PFSceneRef s, uint32_t render_target_id, const PFTransform2F *pattern_transform,
                                             const PFFilter *filter) {
    if (!s || render_target_id >= s->render_targets.size() || s->paints.size() >= 65535) {
        pf::set_last_error("PFScenePushPaintRenderTargetPattern: unknown render target");
        return 0xffff;
    }
    if (filter && filter->kind != PF_FILTER_NONE && filter->kind != PF_FILTER_TEXT && filter->kind != PF_FILTER_BLUR &&
        filter->kind != PF_FILTER_COLOR_MATRIX) {
        pf::set_last_error("PFScenePushPaintRenderTargetPattern: a pattern filter is Text, Blur or ColorMatrix (effects.rs:63-97)");
        return 0xffff;
    }
    PatternOverlay overlay;
    overlay.kind = PatternOverlay::RENDER_TARGET;
    overlay.render_target = render_target_id;
    if (pattern_transform) {
        overlay.transform.m11 = pattern_transform->matrix.m00, overlay.transform.m12 = pattern_transform->matrix.m01;
        overlay.transform.m21 = pattern_transform->matrix.m10, overlay.transform.m22 = pattern_transform->matrix.m11;
        overlay.transform.tx = pattern_transform->vector.x, overlay.transform.ty = pattern_transform->vector.y;
    }
    memset(&overlay.filter, 0, sizeof(overlay.filter));
    if (filter) overlay.filter = *filter;
    const uint16_t id = (uint16_t)s->paints.size();
    s->paints.push_back(PFColorU{255, 255, 255, 255}); // Paint::from_pattern: base colour white (paint.rs:138-146)
    s->overlays.emplace(id, overlay);                    // (pattern paints are never deduplicated)
    s->epoch++;
    return id;
}

uint16_t PFScenePushPaintImagePattern(PFSceneRef s, const PFColorU *pixels, int32_t width, int32_t height,
                                      const PFTransform2F *pattern_transform, uint32_t flags, const PFFilter *filter) {
    if (!s || !pixels || width <= 0 || height <= 0 || (int64_t)width * height > (1ll << 26) || s->paints.size() >= 65535) {
        pf::set_last_error("PFScenePushPaintImagePattern: bad image");
        return 0xffff;
    }
    if (filter && filter->kind != PF_FILTER_NONE && filter->kind != PF_FILTER_TEXT && filter->kind != PF_FILTER_BLUR &&
        filter->kind != PF_FILTER_COLOR_MATRIX) {
        pf::set_last_error("PFScenePushPaintImagePattern: a pattern filter is Text, Blur or ColorMatrix (effects.rs:63-97)");
        return 0xffff;
    }
    PatternOverlay overlay;
    overlay.kind = PatternOverlay::IMAGE;
    overlay.pixels.assign(pixels, pixels + (size_t)width * height);
    overlay.width = width, overlay.height = height;
    overlay.pattern_flags = flags;
    if (pattern_transform) {
        overlay.transform.m11 = pattern_transform->matrix.m00, overlay.transform.m12 = pattern_transform->matrix.m01;
        overlay.transform.m21 = pattern_transform->matrix.m10, overlay.transform.m22 = pattern_transform->matrix.m11;
        overlay.transform.tx = pattern_transform->vector.x, overlay.transform.ty = pattern_transform->vector.y;
    }
    memset(&overlay.filter, 0, sizeof(overlay.filter));
    if (filter) overlay.filter = *filter;
    const uint16_t id = (uint16_t)s->paints.size();
    s->paints.push_back(PFColorU{255, 255, 255, 255}); // Paint::from_pattern: base colour white (paint.rs:138-146)
    s->overlays.emplace(id, std::move(overlay));
    s->epoch++;
    return id;
}

uint16_t PFScenePushPaintGradient(PFSceneRef s, const PFGradient *g) {
    if (!s || !g || (g->stop_count && !g->stops) || g->stop_count > 4096 || s->paints.size() >= 65535 ||
        (g->kind != PF_GRADIENT_LINEAR && g->kind != PF_GRADIENT_RADIAL) ||
        (g->wrap != PF_GRADIENT_WRAP_CLAMP && g->wrap != PF_GRADIENT_WRAP_REPEAT)) {
        pf::set_last_error("PFScenePushPaintGradient: bad gradient");
        return 0xffff;
    }
    for (size_t i = 0; i + 1 < g->stop_count; i++) {
        if (!(g->stops[i].offset <= g->stops[i + 1].offset)) { // Gradient::add_color_stop keeps them sorted (gradient.rs:141-150)
            pf::set_last_error("PFScenePushPaintGradient: colour stops must be sorted by offset");
            return 0xffff;
        }
    }
    PatternOverlay overlay;
    overlay.kind = PatternOverlay::GRADIENT;
    overlay.gradient_kind = g->kind, overlay.gradient_wrap = g->wrap;
    overlay.from[0] = g->from.x, overlay.from[1] = g->from.y, overlay.to[0] = g->to.x, overlay.to[1] = g->to.y;
    overlay.radii[0] = g->radii[0], overlay.radii[1] = g->radii[1];
    overlay.transform.m11 = g->transform.matrix.m00, overlay.transform.m12 = g->transform.matrix.m01;
    overlay.transform.m21 = g->transform.matrix.m10, overlay.transform.m22 = g->transform.matrix.m11;
    overlay.transform.tx = g->transform.vector.x, overlay.transform.ty = g->transform.vector.y;
    overlay.stops.assign(g->stops, g->stops + g->stop_count);
    memset(&overlay.filter, 0, sizeof(overlay.filter));
    const uint16_t id = (uint16_t)s->paints.size();
    s->paints.push_back(PFColorU{255, 255, 255, 255}); // Paint::from_gradient: base colour white (paint.rs:127-136)
    s->overlays.emplace(id, std::move(overlay));
    s->epoch++;
    return id;
}

uint32_t PFSceneGetDrawPathCount(PFSceneRef s) { return (uint32_t)s->draw_paths.size(); }
uint32_t PFSceneGetEpoch(PFSceneRef s) { return s->epoch; }

PFRenderTransformRef PFRenderTransformCreate2D(const PFTransform2F *t) {
    PFRenderTransform *r = new PFRenderTransform();
    // Row-major C struct (c/src/lib.rs:151-163): m00 = m11(), m01 = m12(), m10 = m21(), m11 = m22().
    r->t.m11 = t->matrix.m00, r->t.m12 = t->matrix.m01, r->t.m21 = t->matrix.m10, r->t.m22 = t->matrix.m11;
    r->t.tx = t->vector.x, r->t.ty = t->vector.y;
    return r;
}
void PFRenderTransformDestroy(PFRenderTransformRef t) { delete t; }
PFBuildOptionsRef PFBuildOptionsCreate(void) { return new PFBuildOptions(); }
PFBuildOptionsRef PFBuildOptionsClone(PFBuildOptionsRef o) { return o ? new PFBuildOptions(*o) : nullptr; }
void PFBuildOptionsDestroy(PFBuildOptionsRef o) { delete o; }
void PFBuildOptionsSetTransform(PFBuildOptionsRef o, PFRenderTransformRef t) {
    o->transform = t->t;
    delete t; // consumed (c/src/lib.rs:768-772)
}
void PFBuildOptionsSetDilation(PFBuildOptionsRef o, const PFVector2F *d) {
    o->dilation[0] = d->x;
    o->dilation[1] = d->y;
}
void PFBuildOptionsSetSubpixelAAEnabled(PFBuildOptionsRef o, int32_t enabled) { o->subpixel_aa_enabled = enabled != 0; }

PFCudaStatus PFSceneBuild(PFSceneRef s, PFBuildOptionsRef opts, PFSceneSinkState *sink,
                          PFRenderCommandListenerFn listener, void *userdata) {
    if (!s || !opts || !sink || !listener) {
        pf::set_last_error("PFSceneBuild: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    auto start_time = std::chrono::steady_clock::now();
    PFCudaStatus st;
#define SEND(cmd)                                  \
    do {                                           \
        st = listener(&(cmd), userdata);           \
        if (st != PF_CUDA_OK) return st;           \
    } while (0)

    if (opts->subpixel_aa_enabled) {
        // The reference's GPU prepare mode silently ignores it; it needs the 3x-wide
        // render target and the text filter of the 'next' row f3. Refuse rather than render differently from
        // the CPU tiler.
        pf::set_last_error("subpixel AA is not implemented on the D3D11-level path yet");
        return PF_CUDA_ERROR_UNSUPPORTED;
    }
    // A dilation is applied on the host, after the transform, like the CPU tiler does (prepare_paths above).
    const bool prepared = opts->dilation[0] != 0.0f || opts->dilation[1] != 0.0f;

    // builder.rs:160-164
    PFRenderCommand start = make_command(PF_RENDER_COMMAND_START);
    start.u.start.path_count = s->clip_paths.size() + s->draw_paths.size();
    // needs_readable_framebuffer (builder.rs:372-393): a path outside any render target blends with a mode the
    // shader evaluates itself. (This renderer keeps the pixels of a tile in shared memory, so it needs no copy.)
    start.u.start.needs_readable_framebuffer = 0;
    if (s->any_blend) {
        int depth = 0;
        for (const DisplayItem &item : s->display_list) {
            if (item.kind == DisplayItem::PUSH_RENDER_TARGET) depth++;
            else if (item.kind == DisplayItem::POP_RENDER_TARGET) depth--;
            else if (depth == 0)
                for (uint32_t i = item.a; i < item.b; i++)
                    if (s->draw_paths[i].blend_mode >= PF_BLEND_MODE_DARKEN) start.u.start.needs_readable_framebuffer = 1;
        }
    }
    SEND(start);

    // Paint data (builder.rs:174-180): one TextureMetadataEntry per paint (paint.rs:641-659).
    // Paints are append-only, so (scene id, paint count) identifies the table.
    uint64_t paint_key = mix_key(mix_key(0x9a1f7u, s->id), s->paints.size());
    if (!s->overlays.empty() || s->any_blend) { // texture transforms depend on the build transform (paint.rs:637)
        uint32_t bits[6];
        const float f[6] = {opts->transform.m11, opts->transform.m21, opts->transform.m12, opts->transform.m22,
                            opts->transform.tx,  opts->transform.ty};
        memcpy(bits, f, sizeof(bits));
        for (uint32_t v : bits) paint_key = mix_key(paint_key, v);
        paint_key = mix_key(paint_key, s->epoch);
    }
    // Render targets first (Palette::build_paint_info, paint.rs:399-437: AllocateTexturePage and DeclareRenderTarget
    // precede the metadata). One page per render target, the target covering the whole page; page ids = target ids.
    for (size_t i = 0; i < s->render_targets.size(); i++) {
        PFRenderCommand page = make_command(PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE);
        page.u.allocate_texture_page.page_id = (uint32_t)i;
        page.u.allocate_texture_page.size = PFVector2I{s->render_targets[i].width, s->render_targets[i].height};
        SEND(page);
        PFRenderCommand declare = make_command(PF_RENDER_COMMAND_DECLARE_RENDER_TARGET);
        declare.u.declare_render_target.render_target_id = (uint32_t)i;
        declare.u.declare_render_target.location =
            PFTextureLocation{(uint32_t)i, PFRectI{{0, 0}, {s->render_targets[i].width, s->render_targets[i].height}}};
        SEND(declare);
    }
    // Gradients and images (Palette::assign_paint_locations, paint.rs:456-595; GradientTileBuilder, :813-873): every
    // gradient is one row of a 256 x 256 tile on a page of its own kind, sampled at t = (x + 0.5) / 256; every image
    // gets a page of its own, with a texel of border on the sides it does not repeat on. Pages follow the render
    // targets' pages. (No atlas allocator and no image cache: pages are named and filled afresh by every build.)
    {
        uint32_t next_page = (uint32_t)s->render_targets.size();
        uint32_t n_gradients = 0;
        s->gradient_tiles.clear();
        std::vector<uint16_t> ids;
        for (auto &kv : s->overlays) ids.push_back(kv.first);
        std::sort(ids.begin(), ids.end()); // paint order, like the reference's loop over the palette
        uint32_t gradient_first_page = next_page;
        for (uint16_t id : ids) {
            PatternOverlay &o = s->overlays[id];
            if (o.kind != PatternOverlay::GRADIENT) continue;
            if (n_gradients % 256 == 0) s->gradient_tiles.emplace_back(256 * 256, PFColorU{0, 0, 0, 255}); // ColorU::black()
            o.page = gradient_first_page + n_gradients / 256;
            o.row = (int32_t)(n_gradients % 256);
            o.page_w = o.page_h = 256;
            PFColorU *row = s->gradient_tiles.back().data() + (size_t)o.row * 256;
            for (int x = 0; x < 256; x++) row[x] = gradient_sample(o.stops, ((float)x + 0.5f) / 256.0f);
            n_gradients++;
        }
        next_page += (uint32_t)s->gradient_tiles.size();
        for (size_t t = 0; t < s->gradient_tiles.size(); t++) {
            PFRenderCommand page = make_command(PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE);
            page.u.allocate_texture_page.page_id = gradient_first_page + (uint32_t)t;
            page.u.allocate_texture_page.size = PFVector2I{256, 256};
            SEND(page);
            PFRenderCommand up = make_command(PF_RENDER_COMMAND_UPLOAD_TEXEL_DATA);
            up.u.upload_texel_data.texels = s->gradient_tiles[t].data();
            up.u.upload_texel_data.texel_count = 256 * 256;
            up.u.upload_texel_data.location = PFTextureLocation{gradient_first_page + (uint32_t)t, PFRectI{{0, 0}, {256, 256}}};
            SEND(up);
        }
        for (uint16_t id : ids) {
            PatternOverlay &o = s->overlays[id];
            if (o.kind != PatternOverlay::IMAGE) continue;
            const int bx = (o.pattern_flags & PF_PATTERN_FLAG_REPEAT_X) ? 0 : 1, by = (o.pattern_flags & PF_PATTERN_FLAG_REPEAT_Y) ? 0 : 1;
            o.page = next_page++;
            o.page_w = o.width + 2 * bx, o.page_h = o.height + 2 * by;
            PFRenderCommand page = make_command(PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE);
            page.u.allocate_texture_page.page_id = o.page;
            page.u.allocate_texture_page.size = PFVector2I{o.page_w, o.page_h};
            SEND(page);
            PFRenderCommand up = make_command(PF_RENDER_COMMAND_UPLOAD_TEXEL_DATA);
            up.u.upload_texel_data.texels = o.pixels.data();
            up.u.upload_texel_data.texel_count = o.pixels.size();
            up.u.upload_texel_data.location = PFTextureLocation{o.page, PFRectI{{bx, by}, {bx + o.width, by + o.height}}}; // rect.contract(border)
            SEND(up);
        }
    }
    if (s->built_paint_key != paint_key) {
        s->texture_metadata.resize(s->paints.size());
        // render_transform = the 2-D build transform inverted (builder.rs:168-171)
        const Transform render_transform = transform_inverse(opts->transform);
        for (size_t i = 0; i < s->paints.size(); i++) {
            PFTextureMetadataEntry &e = s->texture_metadata[i];
            memset(&e, 0, sizeof(e));
            e.color_0_transform.matrix = PFMatrix2x2F{1, 0, 0, 1};
            e.base_color = s->paints[i];
            e.color_0_combine_mode = PF_COLOR_COMBINE_MODE_NONE;
            e.blend_mode = PF_BLEND_MODE_SRC_OVER;
            e.filter.kind = PF_FILTER_NONE;
            auto overlay = s->overlays.find((uint16_t)i);
            if (overlay == s->overlays.end()) continue;
            const PatternOverlay &o = overlay->second;
            Transform t; // calculate_texture_transforms (paint.rs:597-639), before `*= render_transform`
            e.filter = o.filter;
            if (o.kind == PatternOverlay::RENDER_TARGET) {
                //   translate(rect_to_uv(rect, scale).lower_left()) * scale(scale * (1, -1)) * pattern.transform().inverse()
                // — v runs bottom-up over a render target (the GL convention, see pf_cuda.h).
                const RenderTargetDesc &rt = s->render_targets[o.render_target];
                const float sx = 1.0f / (float)rt.width, sy = 1.0f / (float)rt.height;
                const Transform inv = transform_inverse(o.transform);
                t.m11 = sx * inv.m11, t.m12 = sx * inv.m12, t.m21 = -sy * inv.m21, t.m22 = -sy * inv.m22;
                t.tx = sx * inv.tx, t.ty = 1.0f - sy * inv.ty;
            } else if (o.kind == PatternOverlay::IMAGE) {
                //   from_scale(texture_scale).translate(rect_to_uv(rect, scale).origin()) * pattern.transform().inverse()
                // The rect is the page (origin 0): the border is NOT added back (the `transform` assign_paint_locations
                // stored is overwritten here), so an image that does not repeat sits one texel further right / down.
                const float sx = 1.0f / (float)o.page_w, sy = 1.0f / (float)o.page_h;
                const Transform inv = transform_inverse(o.transform);
                t.m11 = sx * inv.m11, t.m12 = sx * inv.m12, t.m21 = sy * inv.m21, t.m22 = sy * inv.m22;
                t.tx = sx * inv.tx, t.ty = sy * inv.ty;
            } else if (o.gradient_kind == PF_GRADIENT_LINEAR) {
                // Project the gradient line onto (0..1, v0): v0 = the row's centre (paint.rs:611-618).
                const float v0 = ((float)o.row + 0.5f) * (1.0f / 256.0f);
                const float dx = o.to[0] - o.from[0], dy = o.to[1] - o.from[1];
                const float len2 = dx * dx + dy * dy;
                const float m0x = dx / len2, m0y = dy / len2;
                t.m11 = m0x, t.m12 = m0y, t.m21 = 0.0f, t.m22 = 0.0f;
                t.tx = m0x * -o.from[0] + m0y * -o.from[1], t.ty = v0;
            } else {
                // Radial: the gradient's own transform inverted; the filter finds t and samples (uv_origin + (t, 0))
                // (paint.rs:619-622,788-793: uv_origin = the row's rect contracted by half a texel vertically).
                t = transform_inverse(o.transform);
                e.filter.kind = PF_FILTER_RADIAL_GRADIENT;
                e.filter.flags = 0;
                const float params[8] = {o.from[0], o.from[1], o.to[0], o.to[1], o.radii[0], o.radii[1], 0.0f,
                                         ((float)o.row + 0.5f) * (1.0f / 256.0f)};
                memset(e.filter.params, 0, sizeof(e.filter.params));
                memcpy(e.filter.params, params, sizeof(params));
            }
            const Transform full = transform_mul(t, render_transform);
            e.color_0_transform.matrix = PFMatrix2x2F{full.m11, full.m12, full.m21, full.m22};
            e.color_0_transform.vector = PFVector2F{full.tx, full.ty};
            e.color_0_combine_mode = PF_COLOR_COMBINE_MODE_SRC_IN; // create_texture_metadata (paint.rs:649-653)
        }
        // Blend modes: the reference's TextureMetadataEntry carries one (gpu_data.rs:343) that its palette never sets
        // (paint.rs:576 "FIXME"); here every (paint, blend mode) pair a draw path uses gets an entry of its own after
        // the palette's, and the path's tiles name that entry.
        s->blend_entries.clear();
        if (s->any_blend) {
            for (const Path &dp : s->draw_paths) {
                if (dp.blend_mode == PF_BLEND_MODE_SRC_OVER) continue;
                const uint32_t key = (uint32_t)dp.paint | ((uint32_t)dp.blend_mode << 16);
                if (s->blend_entries.count(key)) continue;
                if (s->texture_metadata.size() >= 65535) {
                    pf::set_last_error("more than 65535 (paint, blend mode) pairs");
                    return PF_CUDA_ERROR_UNSUPPORTED;
                }
                PFTextureMetadataEntry e = s->texture_metadata[dp.paint];
                e.blend_mode = dp.blend_mode;
                s->blend_entries.emplace(key, (uint16_t)s->texture_metadata.size());
                s->texture_metadata.push_back(e);
            }
        }
        s->built_paint_key = paint_key;
    }
    PFRenderCommand meta = make_command(PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA);
    meta.u.upload_texture_metadata.entries = s->texture_metadata.data();
    meta.u.upload_texture_metadata.entry_count = s->texture_metadata.size();
    meta.u.upload_texture_metadata.content_key = paint_key;
    SEND(meta);

    // Scene upload when dirty (builder.rs:190-216). The segment arrays are rebuilt when the scene changed or, with
    // a dilation, when the options that shaped the prepared points did.
    uint64_t segments_key = mix_key(mix_key(0x5e65u, s->id), s->epoch);
    if (prepared) {
        uint32_t bits[8];
        const float f[8] = {opts->transform.m11, opts->transform.m21, opts->transform.m12, opts->transform.m22,
                            opts->transform.tx,  opts->transform.ty,  opts->dilation[0],   opts->dilation[1]};
        memcpy(bits, f, sizeof(bits));
        for (uint32_t v : bits) segments_key = mix_key(segments_key, v);
    }
    // A renderer that owns a strip of tile rows (PFCudaRendererSetStrip / GatherInit; handed down by
    // PFSceneBuildAndRenderCuda) drops every path without a tile in its rows: the builder then neither copies their
    // segments nor builds their records, so at N ranks the host work and the upload per rank shrink with the strip
    // like the device work does. Path ids stay global (DiceMetadataD3D11.global_path_id), so draw order and the
    // z-buffer are those of the whole scene. (Scenes with a display list are built whole.)
    bool general_scene = !s->render_targets.empty() || !s->overlays.empty() || s->any_blend;
    for (const DisplayItem &item : s->display_list) general_scene |= item.kind != DisplayItem::DRAW_PATHS;
    const bool strip_on = pf::g_scene_strip[1] > pf::g_scene_strip[0] && !general_scene;
    const int32_t strip_y0 = pf::g_scene_strip[0], strip_y1 = pf::g_scene_strip[1];
    if (strip_on) {
        uint32_t bits[10];
        const float f[10] = {opts->transform.m11, opts->transform.m21, opts->transform.m12, opts->transform.m22,
                             opts->transform.tx,  opts->transform.ty,  s->view_box.min_x,   s->view_box.min_y,
                             s->view_box.max_x,   s->view_box.max_y};
        memcpy(bits, f, sizeof(bits));
        for (uint32_t v : bits) segments_key = mix_key(segments_key, v);
        segments_key = mix_key(mix_key(segments_key, (uint32_t)strip_y0), (uint32_t)strip_y1);
    }
    if (segments_key == 0) segments_key = 1;
    pf::LapTimer seg_laps;
    seg_laps.lap("paints, keys");
    if (s->segments_key != segments_key) {
        if (prepared) {
            s->prepared_points.resize(s->points.size());
            prepare_paths(s, s->draw_paths, opts->transform, opts->dilation, s->prepared_draw_bounds);
            prepare_paths(s, s->clip_paths, opts->transform, opts->dilation, s->prepared_clip_bounds);
        }
        s->strip_ids.clear();
        if (strip_on) {
            // the rows of prepare_draw_path_for_gpu_binning's tile rect (builder.rs:1075-1079), as in pass 1 below.
            // (At N ranks this pass reads all the path records on a rank's share of the cores: it is kept to a few
            // nanoseconds per path — invariants hoisted, floor / ceil without a libm call, ids compacted per chunk.)
            const size_t n = s->draw_paths.size();
            s->strip_included.assign(n, 0);
            const Transform identity_xf;
            const Transform sxf = prepared ? identity_xf : opts->transform;
            const bool plain = !prepared && sxf.is_identity();
            const RectF view_box = s->view_box;
            // (not paths[i].bounds: the scan streams 16 bytes a path instead of the 40-byte records)
            const RectF *bounds = prepared ? s->prepared_draw_bounds.data() : s->draw_bounds.data();
            const bool transformed = !prepared && !plain;
            uint8_t *included = s->strip_included.data();
            // (int32_t)floorf(v) and (int32_t)ceilf(v) for the finite, in-range values a rect clipped to the view box has
            auto floor_i = [](float v) { const int32_t t = (int32_t)v; return t - ((float)t > v ? 1 : 0); };
            auto ceil_i = [](float v) { const int32_t t = (int32_t)v; return t + ((float)t < v ? 1 : 0); };
            const float strip_lo_px = (float)strip_y0 * 16.0f, strip_lo_px_end = (float)strip_y1 * 16.0f;
            const size_t chunks = pf::chunk_count(n, 8192);
            std::vector<std::vector<uint32_t>> chunk_ids(chunks ? chunks : 1);
            pf::parallel_chunks(n, chunks, [&](size_t c, size_t begin, size_t end) {
                std::vector<uint32_t> &ids = chunk_ids[c];
                ids.resize(end - begin + 1);
                uint32_t *out = ids.data();
                // Most paths lie wholly above or below the strip: floor(max(b.min_y, .) / 16) >= strip_y1 and
                // ceil(min(b.max_y, .) / 16) <= strip_y0 respectively (x / 16 is exact), whatever the view box does.
                // Which side a path of a shuffled scene lies on is a coin toss, so the candidates are compacted
                // WITHOUT a branch (a mispredicted one costs more than the rest of the iteration: 13 -> 3 ns a path);
                // the exact test below then runs over the few that are left, in place.
                size_t candidates = 0;
                if (transformed) {
                    for (size_t i = begin; i < end; i++) {
                        const RectF b = sxf.apply_rect(bounds[i]);
                        out[candidates] = (uint32_t)i;
                        candidates += (size_t)((b.min_y < strip_lo_px_end) & (b.max_y > strip_lo_px));
                    }
                } else {
                    for (size_t i = begin; i < end; i++) {
                        out[candidates] = (uint32_t)i;
                        candidates += (size_t)((bounds[i].min_y < strip_lo_px_end) & (bounds[i].max_y > strip_lo_px));
                    }
                }
                size_t kept = 0;
                for (size_t k = 0; k < candidates; k++) {
                    const uint32_t i = out[k];
                    const RectF b = transformed ? sxf.apply_rect(bounds[i]) : bounds[i];
                    RectF clipped;
                    if (!rect_intersection(b, view_box, clipped)) continue;
                    const float t = 1.0f / 16.0f;
                    const int32_t ty0 = floor_i(clipped.min_y * t), ty1 = ceil_i(clipped.max_y * t);
                    if (ty0 < strip_y1 && ty1 > strip_y0) {
                        included[i] = 1;
                        out[kept++] = i;
                    }
                }
                ids.resize(kept);
            });
            for (const std::vector<uint32_t> &ids : chunk_ids) s->strip_ids.insert(s->strip_ids.end(), ids.begin(), ids.end());
        }
        seg_laps.lap("strip inclusion");
        build_segments(s, prepared ? s->prepared_points.data() : s->points.data(), strip_on ? &s->strip_ids : nullptr);
        seg_laps.lap("segment arrays");
        s->segments_key = segments_key;
        s->upload_serial++;
    }
    const bool dirty = !sink->has_last_scene || sink->last_scene_id != s->id || sink->last_scene_epoch != s->upload_serial;
    if (dirty) {
        PFRenderCommand up = make_command(PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11);
        up.u.upload_scene_d3d11.draw_segments =
            PFSegmentsD3D11{s->seg_points.ptr, s->seg_point_count, s->seg_indices.ptr, s->seg_index_count};
        up.u.upload_scene_d3d11.clip_segments = PFSegmentsD3D11{s->clip_seg_points.ptr, s->clip_seg_point_count,
                                                                s->clip_seg_indices.ptr, s->clip_seg_index_count};
        up.u.upload_scene_d3d11.payload_persists = pf::g_scene_payload_persists ? 1 : 0;
        SEND(up);
        sink->has_last_scene = 1;
        sink->last_scene_id = s->id;
        sink->last_scene_epoch = s->upload_serial;
    }

    bool display_ops = false; // a push / pop anywhere in the display list
    for (const DisplayItem &item : s->display_list) display_ops |= item.kind != DisplayItem::DRAW_PATHS;
    if (display_ops || !s->render_targets.empty() || !s->overlays.empty() || s->any_blend) {
        // A scene with a display list: one batch per DrawPaths item, split again wherever the colour texture
        // changes (build_tile_batches / build_tile_batches_for_draw_path_display_item / fixup_batch_for_new_path_if_possible,
        // builder.rs:327-357,886-940,1227-1243). Built sequentially and afresh every frame (content_key = 0): these
        // scenes are a handful of batches around a text page, not the 100k-path case the cached builder below serves.
        const Transform identity_transform;
        const Transform &gxf = prepared ? identity_transform : opts->transform;
        s->general_batches.clear();
        size_t n_batches = 0;
        for (const DisplayItem &item : s->display_list) n_batches += item.kind == DisplayItem::DRAW_PATHS ? (item.b - item.a) : 0;
        s->general_batches.reserve(n_batches + 1); // (upper bound: pointers into the vector stay valid)
        uint32_t next_batch_id = 32; // MAX_CLIP_BATCHES
        // Clip paths (one level, outside render targets): the clip batch is built once and prepared before the first
        // draw batch that uses it, as in the single-batch build below.
        bool any_clip = false, clip_batch_sent = false;
        for (const Path &dp : s->draw_paths) any_clip |= dp.clip_path != PF_CLIP_PATH_NONE;
        if (any_clip) {
            st = build_clip_batch(s, prepared, gxf, s->view_box);
            if (st != PF_CUDA_OK) return st;
        }
        auto send_batch = [&](GeneralBatch &gb) -> PFCudaStatus {
            if (gb.propagate_metadata.empty()) return PF_CUDA_OK;
            if (gb.clipped_paths && !clip_batch_sent) {
                const PFCudaStatus clip_status = send_clip_batch(s, gxf, listener, userdata);
                if (clip_status != PF_CUDA_OK) return clip_status;
                clip_batch_sent = true;
            }
            PFRenderCommand draw = make_command(PF_RENDER_COMMAND_DRAW_TILES_D3D11);
            PFTileBatchDataD3D11 &b = draw.u.draw_tiles_d3d11.tile_batch_data;
            b.batch_id = next_batch_id++;
            b.path_count = (uint32_t)gb.propagate_metadata.size();
            b.tile_count = gb.tile_count;
            b.segment_count = gb.segment_count;
            b.prepare_info.backdrops = nullptr;
            b.prepare_info.backdrop_count = 0;
            b.prepare_info.propagate_metadata = gb.propagate_metadata.data();
            b.prepare_info.dice_metadata = gb.dice_metadata.data();
            b.prepare_info.tile_path_info = gb.tile_path_info.data();
            b.prepare_info.transform.matrix = PFMatrix2x2F{gxf.m11, gxf.m12, gxf.m21, gxf.m22};
            b.prepare_info.transform.vector = PFVector2F{gxf.tx, gxf.ty};
            b.path_source = PF_PATH_SOURCE_DRAW;
            b.has_clipped_path_info = gb.clipped_paths ? 1 : 0;
            b.clipped_path_info = PFClippedPathInfo{0, gb.clipped_paths, gb.clipped_tiles};
            b.content_key = 0;
            draw.u.draw_tiles_d3d11.has_color_texture = gb.has_color_texture ? 1 : 0;
            draw.u.draw_tiles_d3d11.color_texture = gb.color_texture;
            return listener(&draw, userdata);
        };
        int nesting = 0;
        for (const DisplayItem &item : s->display_list) {
            if (item.kind == DisplayItem::PUSH_RENDER_TARGET) {
                PFRenderCommand push = make_command(PF_RENDER_COMMAND_PUSH_RENDER_TARGET);
                push.u.push_render_target.render_target_id = item.a;
                SEND(push);
                nesting++;
                continue;
            }
            if (item.kind == DisplayItem::POP_RENDER_TARGET) {
                if (nesting == 0) {
                    pf::set_last_error("PFScenePopRenderTarget without a matching push");
                    return PF_CUDA_ERROR_PROTOCOL;
                }
                PFRenderCommand pop = make_command(PF_RENDER_COMMAND_POP_RENDER_TARGET);
                SEND(pop);
                nesting--;
                continue;
            }
            s->general_batches.emplace_back();
            GeneralBatch *gb = &s->general_batches.back();
            for (uint32_t i = item.a; i < item.b; i++) {
                const Path &p = s->draw_paths[i];
                if (p.blend_mode > PF_BLEND_MODE_LUMINOSITY) {
                    pf::set_last_error("unknown blend mode");
                    return PF_CUDA_ERROR_INVALID_ARGUMENT;
                }
                if (p.clip_path != PF_CLIP_PATH_NONE && nesting > 0) {
                    pf::set_last_error("clipped paths inside a render target are not implemented");
                    return PF_CUDA_ERROR_UNSUPPORTED;
                }
                // prepare_draw_path_for_gpu_binning (builder.rs:1058-1095)
                RectF path_bounds = prepared ? s->prepared_draw_bounds[i] : gxf.is_identity() ? p.bounds : gxf.apply_rect(p.bounds);
                RectF clipped;
                const bool has_outline = p.first_contour != p.end_contour;
                if (!((has_outline || !prepared) && rect_intersection(path_bounds, s->view_box, clipped))) continue;
                // BlendMode::is_destructive (effects.rs:222-235): the path's tile map covers the whole view box
                // (BuiltPath::new, builder.rs:430-434). Its empty tiles are still never drawn (builder.rs:1014-1016,
                // propagate.cs.glsl:210-213), so the mode acts on the tiles the path reaches.
                const bool destructive = p.blend_mode == PF_BLEND_MODE_CLEAR || p.blend_mode == PF_BLEND_MODE_COPY ||
                                         p.blend_mode == PF_BLEND_MODE_SRC_IN || p.blend_mode == PF_BLEND_MODE_DEST_IN ||
                                         p.blend_mode == PF_BLEND_MODE_SRC_OUT || p.blend_mode == PF_BLEND_MODE_DEST_ATOP;
                if (destructive) clipped = s->view_box;
                const float k = 1.0f / 16.0f;
                PFRectI tile_rect;
                tile_rect.origin.x = (int32_t)floorf(clipped.min_x * k);
                tile_rect.origin.y = (int32_t)floorf(clipped.min_y * k);
                tile_rect.lower_right.x = (int32_t)ceilf(clipped.max_x * k);
                tile_rect.lower_right.y = (int32_t)ceilf(clipped.max_y * k);
                // the path's colour texture (PaintMetadata::tile_batch_texture, paint.rs:802-804)
                auto overlay = s->overlays.find(p.paint);
                const bool has_texture = overlay != s->overlays.end();
                PFTileBatchTexture texture{0u, 0, PF_PAINT_COMPOSITE_OP_SRC_IN};
                if (has_texture) {
                    const PatternOverlay &o = overlay->second;
                    uint8_t flags = 0; // TextureSamplingFlags (paint.rs:470-476,553-563)
                    if (o.kind == PatternOverlay::RENDER_TARGET) {
                        texture.page = o.render_target;
                    } else if (o.kind == PatternOverlay::GRADIENT) {
                        texture.page = o.page;
                        if (o.gradient_wrap == PF_GRADIENT_WRAP_REPEAT) flags |= PF_TEXTURE_SAMPLING_FLAGS_REPEAT_U;
                    } else {
                        texture.page = o.page;
                        if (o.pattern_flags & PF_PATTERN_FLAG_REPEAT_X) flags |= PF_TEXTURE_SAMPLING_FLAGS_REPEAT_U;
                        if (o.pattern_flags & PF_PATTERN_FLAG_REPEAT_Y) flags |= PF_TEXTURE_SAMPLING_FLAGS_REPEAT_V;
                        if (o.pattern_flags & PF_PATTERN_FLAG_NO_SMOOTHING)
                            flags |= PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MIN | PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MAG;
                    }
                    texture.sampling_flags = flags;
                    if (gb->has_color_texture && (gb->color_texture.page != texture.page ||
                                                  gb->color_texture.sampling_flags != texture.sampling_flags)) { // batch break
                        s->general_batches.emplace_back();
                        gb = &s->general_batches.back();
                    }
                    gb->has_color_texture = true;
                    gb->color_texture = texture;
                }
                uint16_t entry = p.paint; // the texture metadata entry the path's tiles name
                if (p.blend_mode != PF_BLEND_MODE_SRC_OVER)
                    entry = s->blend_entries.at((uint32_t)p.paint | ((uint32_t)p.blend_mode << 16));
                const uint32_t w = (uint32_t)(tile_rect.lower_right.x - tile_rect.origin.x),
                               h = (uint32_t)(tile_rect.lower_right.y - tile_rect.origin.y);
                const uint32_t bi = (uint32_t)gb->propagate_metadata.size();
                // BuiltDrawPath::new (builder.rs:80-94): occludes = opaque paint && the mode ignores the backdrop
                // (SrcOver, Clear: effects.rs:202-204); a render-target pattern is never "obviously opaque" (pattern.rs:264-272).
                const bool occludes = !has_texture && s->paints[p.paint].a == 255 &&
                                      (p.blend_mode == PF_BLEND_MODE_SRC_OVER || p.blend_mode == PF_BLEND_MODE_CLEAR);
                PFPropagateMetadataD3D11 pm;
                memset(&pm, 0, sizeof(pm));
                pm.tile_rect = tile_rect;
                pm.tile_offset = gb->tile_count;
                pm.path_index = bi;
                pm.z_write = occludes ? 1 : 0;
                pm.clip_path_index = p.clip_path == PF_CLIP_PATH_NONE ? PF_PATH_INDEX_NONE : s->clip_batch_index[p.clip_path];
                pm.backdrop_offset = gb->column_count;
                if (p.clip_path != PF_CLIP_PATH_NONE) gb->clipped_paths++, gb->clipped_tiles += w * h;
                gb->propagate_metadata.push_back(pm);
                gb->dice_metadata.push_back(PFDiceMetadataD3D11{i, s->draw_segment_ranges[2 * i], gb->segment_count, 0});
                PFTilePathInfoD3D11 tp;
                tp.tile_min_x = (int16_t)tile_rect.origin.x;
                tp.tile_min_y = (int16_t)tile_rect.origin.y;
                tp.tile_max_x = (int16_t)tile_rect.lower_right.x;
                tp.tile_max_y = (int16_t)tile_rect.lower_right.y;
                tp.first_tile_index = gb->tile_count;
                tp.color = entry;
                tp.ctrl = p.fill_rule == PF_FILL_RULE_EVEN_ODD ? PF_TILE_CTRL_MASK_EVEN_ODD : PF_TILE_CTRL_MASK_WINDING;
                tp.backdrop = 0;
                gb->tile_path_info.push_back(tp);
                gb->tile_count += w * h;
                gb->column_count += w;
                gb->segment_count += s->draw_segment_ranges[2 * i + 1] - s->draw_segment_ranges[2 * i];
            }
            // the batches of this display item, in order
            for (GeneralBatch &b : s->general_batches) {
                if (b.tile_path_info.empty() || b.segment_count == 0xffffffffu) continue;
                st = send_batch(b);
                if (st != PF_CUDA_OK) return st;
                b.segment_count = 0xffffffffu; // sent
            }
        }
        if (nesting != 0) {
            pf::set_last_error("a render target is still pushed at the end of the display list");
            return PF_CUDA_ERROR_PROTOCOL;
        }
        PFRenderCommand finish = make_command(PF_RENDER_COMMAND_FINISH);
        finish.u.finish.cpu_build_time_ns =
            (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - start_time).count();
        SEND(finish);
        return PF_CUDA_OK;
    }

    // build_tile_batches at the D3D11 level (builder.rs:327-357, 886-1056): solid colours never
    // break a batch (fixup_batch_for_new_path_if_possible, :1227-1243), so one DrawTilesD3D11.
    // PrepareMode::GPU { transform } (options.rs:165-180); prepared points are already in device space.
    const Transform identity_transform;
    const Transform &xf = prepared ? identity_transform : opts->transform;
    const RectF effective_view_box = s->view_box; // subpixel AA refused above (scene.rs:276-282)
    // The batch arrays depend only on (scene, epoch, transform, view box): keep them across frames
    // and tell the renderer through content_key that nothing changed.
    uint64_t batch_key = mix_key(mix_key(0xb47c4u, s->id), s->epoch);
    {
        uint32_t bits[10];
        const float f[10] = {xf.m11, xf.m21, xf.m12, xf.m22, xf.tx, xf.ty, effective_view_box.min_x,
                             effective_view_box.min_y, effective_view_box.max_x, effective_view_box.max_y};
        memcpy(bits, f, sizeof(bits));
        for (uint32_t v : bits) batch_key = mix_key(batch_key, v);
        if (prepared || strip_on) batch_key = mix_key(batch_key, segments_key);
        if (batch_key == 0) batch_key = 1;
    }
    uint32_t tile_count = s->built_tile_count, segment_count = s->built_segment_count;
    const bool rebuild = s->built_key != batch_key;
    if (rebuild) {
        tile_count = segment_count = 0;
    }
    auto t_loop0 = std::chrono::steady_clock::now();
    pf::LapTimer laps;
    if (rebuild) {
        const size_t n_paths = s->draw_paths.size();
        std::atomic<int> unsupported{0}; // 2: a blend mode other than SrcOver
        {
            const PFCudaStatus clip_status = build_clip_batch(s, prepared, xf, effective_view_box);
            if (clip_status != PF_CUDA_OK) return clip_status;
        }
        // Pass 1 (parallel): prepare_draw_path_for_gpu_binning (builder.rs:1058-1095) — the tile rect
        // of every path; an empty rect marks a path outside the view box (skipped by the builder).
        // Each chunk also sums what its kept paths add to the batch's running offsets, so that pass 2
        // (same chunks) can assign the offsets of TileBatchDataD3D11::push (builder.rs:660-721) in
        // parallel: a two-level scan, integer adds only.
        struct ChunkSums {
            uint32_t kept = 0, tiles = 0, columns = 0, segments = 0;
            uint32_t clipped = 0, clipped_tiles = 0; // kept paths with a clip path, and their tiles
            uint32_t pad[10]; // one cache line per chunk
        };
        s->path_tile_rects.resize(n_paths);
        // (a strip build visits only the paths that reach the strip: `items` of them, ids in strip_ids)
        const uint32_t *ids = strip_on ? s->strip_ids.data() : nullptr;
        const size_t items = strip_on ? s->strip_ids.size() : n_paths;
        const size_t chunks = pf::chunk_count(items, 4096);
        std::vector<ChunkSums> sums(chunks + 1);
        pf::parallel_chunks(items, chunks, [&](size_t chunk, size_t begin, size_t end) {
            ChunkSums sum;
            for (size_t k = begin; k < end; k++) {
                const size_t i = ids ? ids[k] : k;
                const Path &p = s->draw_paths[i];
                if (p.blend_mode != PF_BLEND_MODE_SRC_OVER) unsupported.store(2, std::memory_order_relaxed);
                PFRectI tile_rect{{0, 0}, {0, 0}};
                RectF path_bounds = prepared ? s->prepared_draw_bounds[i]
                                             : xf.is_identity() ? p.bounds : xf.apply_rect(p.bounds);
                RectF clipped;
                // (An outline without contours has the bounds (0, 0, 0, 0); grown by a dilation they would reach
                // into the view box and keep a path that has nothing to tile. The CPU tiler builds one blank tile
                // for it, which never reaches a batch; here the path is skipped like any other invisible one.)
                const bool has_outline = p.first_contour != p.end_contour;
                if ((has_outline || !prepared) && rect_intersection(path_bounds, effective_view_box, clipped)) {
                    // round_rect_out_to_tile_bounds (tiles.rs:64-66); floor/ceil results are integral,
                    // so the float -> int conversion is exact in any rounding mode.
                    const float k = 1.0f / 16.0f;
                    tile_rect.origin.x = (int32_t)floorf(clipped.min_x * k);
                    tile_rect.origin.y = (int32_t)floorf(clipped.min_y * k);
                    tile_rect.lower_right.x = (int32_t)ceilf(clipped.max_x * k);
                    tile_rect.lower_right.y = (int32_t)ceilf(clipped.max_y * k);
                    const uint32_t w = (uint32_t)(tile_rect.lower_right.x - tile_rect.origin.x),
                                   h = (uint32_t)(tile_rect.lower_right.y - tile_rect.origin.y);
                    sum.kept++;
                    if (p.clip_path != PF_CLIP_PATH_NONE) sum.clipped++, sum.clipped_tiles += w * h;
                    sum.tiles += w * h;
                    sum.columns += w;
                    sum.segments += s->draw_segment_ranges[2 * i + 1] - s->draw_segment_ranges[2 * i];
                } else {
                    tile_rect.origin.x = 1, tile_rect.lower_right.x = 0; // "skipped" marker (a kept rect has max >= min)
                }
                s->path_tile_rects[i] = tile_rect;
            }
            sums[chunk + 1] = sum;
        });
        laps.lap("scene pass 1");
        if (unsupported.load() != 0) {
            pf::set_last_error("only BlendMode::SrcOver is on the hot path");
            return PF_CUDA_ERROR_UNSUPPORTED;
        }
        for (size_t c = 1; c <= chunks; c++) { // exclusive prefix: sums[c] = totals of chunks before c
            sums[c].kept += sums[c - 1].kept;
            sums[c].tiles += sums[c - 1].tiles;
            sums[c].columns += sums[c - 1].columns;
            sums[c].segments += sums[c - 1].segments;
            sums[c].clipped += sums[c - 1].clipped;
            sums[c].clipped_tiles += sums[c - 1].clipped_tiles;
        }
        s->built_clipped_path_count = sums[chunks].clipped;
        s->built_clipped_tile_count = sums[chunks].clipped_tiles;
        const uint32_t kept = sums[chunks].kept;
        tile_count = sums[chunks].tiles;
        segment_count = sums[chunks].segments;
        s->propagate_metadata.resize(kept);
        s->dice_metadata.resize(kept);
        s->tile_path_info.resize(kept);
        // Pass 2 (parallel, same chunks): the records.
        pf::parallel_chunks(items, chunks, [&](size_t chunk, size_t begin, size_t end) {
            uint32_t bi = sums[chunk].kept, tile_off = sums[chunk].tiles, col_off = sums[chunk].columns,
                     seg_off = sums[chunk].segments;
            for (size_t k = begin; k < end; k++) {
                const size_t i = ids ? ids[k] : k;
                const PFRectI &tile_rect = s->path_tile_rects[i];
                if (tile_rect.origin.x > tile_rect.lower_right.x) continue;
                const Path &p = s->draw_paths[i];
                // BuiltDrawPath::new (builder.rs:80-94): occludes = opaque paint && SrcOver.
                const bool occludes = s->paints[p.paint].a == 255;
                const uint8_t ctrl = p.fill_rule == PF_FILL_RULE_EVEN_ODD ? PF_TILE_CTRL_MASK_EVEN_ODD : PF_TILE_CTRL_MASK_WINDING;
                PFPropagateMetadataD3D11 pm;
                memset(&pm, 0, sizeof(pm));
                pm.tile_rect = tile_rect;
                pm.tile_offset = tile_off;
                pm.path_index = bi;
                pm.z_write = occludes ? 1 : 0;
                pm.clip_path_index = p.clip_path == PF_CLIP_PATH_NONE ? PF_PATH_INDEX_NONE : s->clip_batch_index[p.clip_path];
                pm.backdrop_offset = col_off;
                s->propagate_metadata[bi] = pm;
                s->dice_metadata[bi] = PFDiceMetadataD3D11{(uint32_t)i, s->draw_segment_ranges[2 * i], seg_off, 0};
                PFTilePathInfoD3D11 tp;
                tp.tile_min_x = (int16_t)tile_rect.origin.x;
                tp.tile_min_y = (int16_t)tile_rect.origin.y;
                tp.tile_max_x = (int16_t)tile_rect.lower_right.x;
                tp.tile_max_y = (int16_t)tile_rect.lower_right.y;
                tp.first_tile_index = tile_off;
                tp.color = p.paint;
                tp.ctrl = ctrl;
                tp.backdrop = 0;
                s->tile_path_info[bi] = tp;
                const uint32_t w = (uint32_t)(tile_rect.lower_right.x - tile_rect.origin.x),
                               h = (uint32_t)(tile_rect.lower_right.y - tile_rect.origin.y);
                bi++;
                tile_off += w * h;
                col_off += w;
                seg_off += s->draw_segment_ranges[2 * i + 1] - s->draw_segment_ranges[2 * i];
            }
        });
    }
    laps.lap("scene pass 2");
    if (getenv("PF_HOST_TIMING"))
        fprintf(stderr, "PFSceneBuild (segment arrays %s): %.3f ms before the path loop, %.3f ms in it\n",
                s->seg_points.pinned ? "pinned" : "pageable",
                std::chrono::duration<double, std::milli>(t_loop0 - start_time).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_loop0).count());
    s->built_key = batch_key;
    s->built_tile_count = tile_count;
    s->built_segment_count = segment_count;
    const bool has_clips = !s->propagate_metadata.empty() && s->built_clipped_path_count > 0;
    if (has_clips) {
        st = send_clip_batch(s, xf, listener, userdata);
        if (st != PF_CUDA_OK) return st;
    }
    if (!s->propagate_metadata.empty()) {
        PFRenderCommand draw = make_command(PF_RENDER_COMMAND_DRAW_TILES_D3D11);
        PFTileBatchDataD3D11 &b = draw.u.draw_tiles_d3d11.tile_batch_data;
        b.batch_id = 32; // MAX_CLIP_BATCHES: draw batch ids start there (builder.rs:50,864)
        b.path_count = (uint32_t)s->propagate_metadata.size();
        b.tile_count = tile_count;
        b.segment_count = segment_count;
        b.prepare_info.backdrops = nullptr; // all-zero initial backdrops (init_backdrops, builder.rs:762-768)
        b.prepare_info.backdrop_count = 0;
        b.prepare_info.propagate_metadata = s->propagate_metadata.data();
        b.prepare_info.dice_metadata = s->dice_metadata.data();
        b.prepare_info.tile_path_info = s->tile_path_info.data();
        b.prepare_info.transform.matrix = PFMatrix2x2F{xf.m11, xf.m12, xf.m21, xf.m22};
        b.prepare_info.transform.vector = PFVector2F{xf.tx, xf.ty};
        b.path_source = PF_PATH_SOURCE_DRAW;
        b.has_clipped_path_info = has_clips ? 1 : 0;
        b.clipped_path_info = PFClippedPathInfo{0, s->built_clipped_path_count, s->built_clipped_tile_count};
        b.content_key = batch_key;
        draw.u.draw_tiles_d3d11.has_color_texture = 0;
        SEND(draw);
    }

    PFRenderCommand finish = make_command(PF_RENDER_COMMAND_FINISH);
    finish.u.finish.cpu_build_time_ns =
        (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - start_time).count();
    SEND(finish);
#undef SEND
    return PF_CUDA_OK;
}

// PFSceneBuild for a renderer that owns the tile rows [tile_y0, tile_y1) of the frame (PFCudaRendererSetStrip): paths
// without a tile in those rows are left out of the segment arrays and of the batch. PFSceneBuildAndRenderCuda does this
// by itself with the renderer's strip.
PFCudaStatus PFSceneBuildForStrip(PFSceneRef s, PFBuildOptionsRef opts, PFSceneSinkState *sink, PFRenderCommandListenerFn listener,
                                  void *userdata, int32_t tile_y0, int32_t tile_y1) {
    pf::g_scene_strip[0] = tile_y0, pf::g_scene_strip[1] = tile_y1;
    const PFCudaStatus st = PFSceneBuild(s, opts, sink, listener, userdata);
    pf::g_scene_strip[0] = pf::g_scene_strip[1] = 0;
    return st;
}

} // extern "C"
