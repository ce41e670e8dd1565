// oracle/pf_oracle.cpp — TEST INFRASTRUCTURE ONLY (see pf_oracle.h).
//
// CPU restatement of the reference's CPU tiler (renderer level D3D9, SequentialExecutor order)
// and of the fill / composite shader math. Every function cites the reference file:line it
// follows (paths relative to /root/reference). PARITY UNPINNED: the reference has no tests or
// golden vectors for this path and cannot be built here.
//
// Arithmetic contract: IEEE binary32, one rounding per operator, no FMA
// contraction. Build with -O2 -ffp-contract=off -msse4.1, never -ffast-math.

#include "pf_oracle.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// SIMD-lane semantics restated as scalars (simd/src/x86/mod.rs).
// ---------------------------------------------------------------------------------------------

// _mm_min_ps(a, b) = a < b ? a : b  (second operand on NaN / equal)   simd/src/x86/mod.rs:247-249
static inline float sse_min(float a, float b) { return a < b ? a : b; }
// _mm_max_ps(a, b) = a > b ? a : b                                     simd/src/x86/mod.rs:252-254
static inline float sse_max(float a, float b) { return a > b ? a : b; }
// F32x4::clamp = self.max(min).min(max)                                simd/src/x86/mod.rs:257-259
static inline float sse_clamp(float x, float lo, float hi) { return sse_min(sse_max(x, lo), hi); }
// F32x4::to_i32x4 = _mm_cvtps_epi32, round-to-nearest-even             simd/src/x86/mod.rs:318-320
static inline int32_t cvtps(float x) { return (int32_t)lrintf(x); }

struct V2 {
    float x, y;
};
static inline V2 v2(float x, float y) { return V2{x, y}; }
static inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
static inline bool operator==(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
static inline bool operator!=(V2 a, V2 b) { return !(a == b); }

struct RectF {
    float min_x, min_y, max_x, max_y;
};
struct RectI {
    int32_t min_x, min_y, max_x, max_y;
    int32_t width() const { return max_x - min_x; }
    int32_t height() const { return max_y - min_y; }
    // geometry/src/rect.rs:398-407
    bool contains_point(int32_t x, int32_t y) const {
        return min_x <= x && min_y <= y && x <= max_x - 1 && y <= max_y - 1;
    }
};

// Transform2F: matrix lanes [m11, m21, m12, m22] + vector. geometry/src/transform2d.rs:123-130,312-318
struct Transform2F {
    float m11, m21, m12, m22, tx, ty;
    bool is_identity() const {
        return m11 == 1.0f && m21 == 0.0f && m12 == 0.0f && m22 == 1.0f && tx == 0.0f && ty == 0.0f;
    }
    // Matrix2x2F * Vector2F: halves = m * v.xxyy; halves.xy + halves.zw; then + vector.
    V2 apply(V2 p) const {
        float hx = m11 * p.x, hy = m21 * p.x, hz = m12 * p.y, hw = m22 * p.y;
        return v2((hx + hz) + tx, (hy + hw) + ty);
    }
};

// Transform2F * Transform2F (geometry/src/transform2d.rs:297-306, Matrix2x2F mul :112-119).
static Transform2F transform_mul(const Transform2F &a, const Transform2F &b) {
    Transform2F r;
    // self.0.xyxy() * other.0.xxzz() + self.0.zwzw() * other.0.yyww()
    r.m11 = a.m11 * b.m11 + a.m12 * b.m21;
    r.m21 = a.m21 * b.m11 + a.m22 * b.m21;
    r.m12 = a.m11 * b.m12 + a.m12 * b.m22;
    r.m22 = a.m21 * b.m12 + a.m22 * b.m22;
    V2 t = a.apply(v2(b.tx, b.ty));
    r.tx = t.x;
    r.ty = t.y;
    return r;
}

// ---------------------------------------------------------------------------------------------
// Outline (content/src/outline.rs), reduced to what the tiler reads.
// ---------------------------------------------------------------------------------------------

struct Contour {
    std::vector<V2> points;
    std::vector<uint8_t> flags;
    RectF bounds{0, 0, 0, 0};
};
struct Outline {
    std::vector<Contour> contours;
    RectF bounds{0, 0, 0, 0};
};

// union_rect / RectF::union_point (outline.rs:1086-1092, rect.rs:109-112). Vector2F::min/max are
// the SSE min/max of (self, other).
static inline void union_point(RectF &b, V2 p, bool first) {
    if (first) {
        b = RectF{p.x, p.y, p.x, p.y};
    } else {
        b.min_x = sse_min(b.min_x, p.x);
        b.min_y = sse_min(b.min_y, p.y);
        b.max_x = sse_max(b.max_x, p.x);
        b.max_y = sse_max(b.max_y, p.y);
    }
}
static inline RectF union_rect(RectF a, RectF b) { // rect.rs:114-120
    return RectF{sse_min(a.min_x, b.min_x), sse_min(a.min_y, b.min_y), sse_max(a.max_x, b.max_x),
                 sse_max(a.max_y, b.max_y)};
}

// Outline bounds as push_contour / transform maintain them (outline.rs:180-192, 208-221,
// 884-896, 921-926): union over non-empty contours of the min/max over ALL their points.
static void recompute_bounds(Outline &o) {
    bool have = false;
    for (Contour &c : o.contours) {
        for (size_t i = 0; i < c.points.size(); i++) union_point(c.bounds, c.points[i], i == 0);
        if (c.points.empty()) {
            // Outline::transform unions contour.bounds even for empty contours, but an Outline
            // never holds an empty contour (push_contour drops them, outline.rs:181-183).
            continue;
        }
        o.bounds = have ? union_rect(o.bounds, c.bounds) : c.bounds;
        have = true;
    }
    if (!have) o.bounds = RectF{0, 0, 0, 0};
}

static Outline load_outline(const PFOScene &s, uint32_t first_contour, uint32_t end_contour) {
    Outline o;
    for (uint32_t ci = first_contour; ci < end_contour; ci++) {
        uint32_t p0 = s.contour_offsets[ci], p1 = s.contour_offsets[ci + 1];
        if (p0 == p1) continue; // push_contour drops empty contours
        Contour c;
        c.points.reserve(p1 - p0);
        c.flags.reserve(p1 - p0);
        for (uint32_t pi = p0; pi < p1; pi++) {
            c.points.push_back(v2(s.points[2 * pi], s.points[2 * pi + 1]));
            c.flags.push_back(s.point_flags[pi]);
        }
        o.contours.push_back(std::move(c));
    }
    recompute_bounds(o);
    return o;
}

// Orientation::from_outline (content/src/orientation.rs:28-49). Vector2F::det = x0*y1 - y0*x1.
static bool outline_is_cw(const Outline &o) {
    float area = 0.0f;
    for (const Contour &c : o.contours) {
        if (c.points.empty()) continue;
        V2 prev = c.points.back();
        for (V2 next : c.points) {
            area += prev.x * next.y - prev.y * next.x;
            prev = next;
        }
    }
    return !(area <= 0.0f); // area <= 0 => Ccw
}

static inline V2 normalize(V2 v) { // geometry/src/vector.rs: self * (1.0 / self.length())
    float len = sqrtf(v.x * v.x + v.y * v.y);
    float inv = 1.0f / len;
    return v2(v.x * inv, v.y * inv);
}

// ContourDilator::dilate (content/src/dilation.rs:34-125).
static void dilate_contour(Contour &c, V2 amount, bool cw) {
    const uint32_t n = (uint32_t)c.points.size();
    if (n == 0) return;
    auto prev_index = [n](uint32_t i) { return i == 0 ? n - 1 : i - 1; };  // outline.rs prev_point_index_of
    auto next_index = [n](uint32_t i) { return i + 1 == n ? 0u : i + 1; }; // next_point_index_of
    V2 scale = cw ? v2(amount.x * -1.0f, amount.y * 1.0f) : v2(amount.x * 1.0f, amount.y * -1.0f);

    V2 first_position = c.points[0];
    uint32_t prev_point_index = 0;
    V2 prev_position;
    for (;;) {
        prev_point_index = prev_index(prev_point_index);
        prev_position = c.points[prev_point_index];
        if (prev_point_index == 0 || prev_position != first_position) break;
    }
    uint32_t first_point_index = next_index(prev_point_index);
    uint32_t current_point_index = first_point_index;
    V2 position = first_position;
    V2 prev_vector = normalize(position - prev_position);
    for (;;) {
        uint32_t next_point_index = current_point_index;
        V2 next_position;
        for (;;) {
            next_point_index = next_index(next_point_index);
            if (next_point_index == first_point_index) {
                next_position = first_position;
                break;
            }
            next_position = c.points[next_point_index];
            if (next_point_index == current_point_index || next_position != position) break;
        }
        V2 next_vector = normalize(next_position - position);
        V2 bisector = v2(prev_vector.y + next_vector.y, prev_vector.x + next_vector.x);
        float bisector_length = sqrtf(bisector.x * bisector.x + bisector.y * bisector.y);
        V2 scaled_bisector;
        if (bisector_length == 0.0f) {
            scaled_bisector = v2(0, 0);
        } else {
            float inv = 1.0f / bisector_length;
            scaled_bisector = v2((bisector.x * scale.x) * inv, (bisector.y * scale.y) * inv);
        }
        V2 new_position = position - scaled_bisector;
        uint32_t point_index = current_point_index;
        while (point_index != next_point_index) {
            c.points[point_index] = new_position;
            point_index = next_index(point_index);
        }
        if (next_point_index == first_point_index) break;
        prev_vector = next_vector;
        position = next_position;
        current_point_index = next_point_index;
    }
}

// Scene::apply_render_options, 2-D branch (renderer/src/scene.rs:228-273).
static Outline apply_render_options(const Outline &original, const PFOBuildOptions &opt) {
    Outline o = original; // clone; close_all_contours only sets the `closed` flag
    bool is_2d = false;
    Transform2F t{1, 0, 0, 1, 0, 0};
    if (opt.has_transform) {
        Transform2F given{opt.transform[0], opt.transform[1], opt.transform[2],
                          opt.transform[3], opt.transform[4], opt.transform[5]};
        if (!given.is_identity()) { // RenderTransform::prepare, options.rs:93-97
            is_2d = true;
            t = given;
        }
    }
    if (is_2d || opt.subpixel_aa_enabled) {
        if (opt.subpixel_aa_enabled) t = transform_mul(t, Transform2F{3, 0, 0, 1, 0, 0});
        if (!t.is_identity()) { // Outline::transform, outline.rs:208-221
            for (Contour &c : o.contours)
                for (V2 &p : c.points) p = t.apply(p);
            recompute_bounds(o);
        }
    }
    if (!(opt.dilation[0] == 0.0f && opt.dilation[1] == 0.0f)) { // outline.rs:243-249
        bool cw = outline_is_cw(o);
        V2 amount = v2(opt.dilation[0], opt.dilation[1]);
        for (Contour &c : o.contours) dilate_contour(c, amount, cw);
        // RectF::dilate (rect.rs:171-175); contour bounds are not touched by the reference.
        o.bounds = RectF{o.bounds.min_x - amount.x, o.bounds.min_y - amount.y,
                         o.bounds.max_x + amount.x, o.bounds.max_y + amount.y};
    }
    return o;
}

// ---------------------------------------------------------------------------------------------
// Built path data (renderer/src/builder.rs:99-122, 414-484).
// ---------------------------------------------------------------------------------------------

constexpr uint32_t INVALID_ALPHA = 0xffffffffu;
constexpr int TILE = 16; // tiles.rs:19-20
constexpr uint8_t TILE_CTRL_MASK_WINDING = 0x1, TILE_CTRL_MASK_EVEN_ODD = 0x2; // gpu_data.rs:32-33

struct BuiltPath {
    RectI tile_bounds{0, 0, 0, 0};
    std::vector<PFOTileObjectPrimitive> tiles; // DenseTileMap, row-major (tile_map.rs:24-33)
    std::vector<int32_t> backdrops;            // per column, rows above the rect
    std::vector<PFOClip> clip_tiles;           // only when the draw path has a clip path
    std::vector<PFOFill> fills;
    std::vector<float> lines;                  // kept only on request (dice parity)
    uint64_t n_lines = 0, n_input_segments = 0;
    uint8_t ctrl = 0;
    uint8_t fill_rule = 0;
    bool occludes = false;
};

struct SceneCtx {
    RectF view_box;            // scene.view_box(): what process_line_segment clips to (quirk 1)
    RectF effective_view_box;  // scene.effective_view_box(): bounds / tile rect
    bool keep_lines = false;
    int32_t strip_y0 = 0, strip_y1 = 0;
    // next_alpha_tile_indices[0], builder.rs:55. On its own cache line: every worker increments it per alpha tile,
    // and every line segment reads the fields above.
    alignas(64) std::atomic<uint32_t> next_alpha_tile{0};
    char pad_[60];
};

// RectF::intersects + intersection (rect.rs:122-137): strict '<' on all four lanes.
static bool rect_intersection(RectF a, RectF b, RectF &out) {
    bool hit = a.min_x < b.max_x && a.min_y < b.max_y && b.min_x < a.max_x && b.min_y < a.max_y;
    if (!hit) return false;
    out = RectF{sse_max(a.min_x, b.min_x), sse_max(a.min_y, b.min_y), sse_min(a.max_x, b.max_x),
                sse_min(a.max_y, b.max_y)};
    return true;
}

// tiles::round_rect_out_to_tile_bounds (tiles.rs:64-66): rect * (1/16) -> round_out -> to_i32.
static RectI round_rect_out_to_tile_bounds(RectF r) {
    const float sx = 1.0f / (float)TILE, sy = 1.0f / (float)TILE;
    return RectI{cvtps(floorf(r.min_x * sx)), cvtps(floorf(r.min_y * sy)), cvtps(ceilf(r.max_x * sx)),
                 cvtps(ceilf(r.max_y * sy))};
}

struct ObjectBuilder {
    BuiltPath *bp;
    SceneCtx *ctx;
    uint32_t path_id;

    // builder.rs:578-592
    bool local_index(int32_t tx, int32_t ty, uint32_t &idx) const {
        const RectI &r = bp->tile_bounds;
        if (!r.contains_point(tx, ty)) return false;
        idx = (uint32_t)((tx - r.min_x) + r.width() * (ty - r.min_y));
        return true;
    }

    // ObjectBuilder::add_fill (builder.rs:509-553)
    void add_fill(V2 from, V2 to, int32_t tx, int32_t ty) {
        uint32_t idx;
        if (!local_index(tx, ty, idx)) return;
        // tile_upper_left = tile_coords.to_f32().xyxy() * 16
        float ulx = (float)tx * 16.0f, uly = (float)ty * 16.0f;
        // (segment - tile_upper_left) * 256, clamp [0, 4095], cvtps
        const float lo = 0.0f, hi = (float)(TILE * 256 - 1);
        int32_t from_x = cvtps(sse_clamp((from.x - ulx) * 256.0f, lo, hi));
        int32_t from_y = cvtps(sse_clamp((from.y - uly) * 256.0f, lo, hi));
        int32_t to_x = cvtps(sse_clamp((to.x - ulx) * 256.0f, lo, hi));
        int32_t to_y = cvtps(sse_clamp((to.y - uly) * 256.0f, lo, hi));
        if (from_x == to_x) return; // cull degenerate fills
        // get_or_allocate_alpha_tile_index (builder.rs:555-576), level 0
        PFOTileObjectPrimitive &tile = bp->tiles[idx];
        if (tile.alpha_tile_id == INVALID_ALPHA)
            tile.alpha_tile_id = ctx->next_alpha_tile.fetch_add(1, std::memory_order_relaxed);
        bp->fills.push_back(PFOFill{(uint16_t)from_x, (uint16_t)from_y, (uint16_t)to_x,
                                    (uint16_t)to_y, tile.alpha_tile_id});
    }

    // ObjectBuilder::adjust_alpha_tile_backdrop (builder.rs:595-616)
    void adjust_backdrop(int32_t tx, int32_t ty, int8_t delta) {
        const RectI &r = bp->tile_bounds;
        int32_t ox = tx - r.min_x, oy = ty - r.min_y;
        if (ox < 0 || ox >= r.width() || oy >= r.height()) return;
        if (oy < 0) {
            bp->backdrops[ox] += (int32_t)delta;
            return;
        }
        PFOTileObjectPrimitive &tile = bp->tiles[(size_t)oy * r.width() + ox];
        tile.backdrop = (int8_t)(uint8_t)((uint8_t)tile.backdrop + (uint8_t)delta); // wrapping i8 add (release build)
    }
};

// util::lerp (geometry/src/util.rs:25-27)
static inline float lerp(float a, float b, float t) { return a + (b - a) * t; }

// clip::clip_line_segment_to_rect (content/src/clip.rs:494-565), Cohen-Sutherland.
static bool clip_line_segment_to_rect(V2 &from, V2 &to, RectF rect) {
    auto outcode = [&](V2 p) -> unsigned {
        unsigned o = 0;
        if (p.x < rect.min_x) o |= 0x01; // LEFT
        if (p.y < rect.min_y) o |= 0x04; // TOP
        if (p.x > rect.max_x) o |= 0x02; // RIGHT
        if (p.y > rect.max_y) o |= 0x08; // BOTTOM
        return o;
    };
    unsigned oc_from = outcode(from), oc_to = outcode(to);
    for (;;) {
        if (oc_from == 0 && oc_to == 0) return true;
        if ((oc_from & oc_to) != 0) return false;
        bool clip_from = oc_from > oc_to;
        V2 point = clip_from ? from : to;
        unsigned oc = clip_from ? oc_from : oc_to;
        if (oc & 0x01) {
            point = v2(rect.min_x, lerp(from.y, to.y, (rect.min_x - from.x) / (to.x - from.x)));
        } else if (oc & 0x02) {
            point = v2(rect.max_x, lerp(from.y, to.y, (rect.max_x - from.x) / (to.x - from.x)));
        } else if (oc & 0x04) {
            point = v2(lerp(from.x, to.x, (rect.min_y - from.y) / (to.y - from.y)), rect.min_y);
        } else if (oc & 0x08) {
            point = v2(lerp(from.x, to.x, (rect.max_y - from.y) / (to.y - from.y)), rect.max_y);
        }
        if (clip_from) {
            from = point;
            oc_from = outcode(point);
        } else {
            to = point;
            oc_to = outcode(point);
        }
    }
}

// process_line_segment (renderer/src/tiler.rs:191-308): Nehab-Hoppe lattice clipping with an
// Amanatides-Woo tile walk.
static void process_line_segment(V2 from, V2 to, ObjectBuilder &ob) {
    SceneCtx &ctx = *ob.ctx;
    ob.bp->n_lines++;
    if (ctx.keep_lines) {
        float l[4] = {from.x, from.y, to.x, to.y};
        ob.bp->lines.insert(ob.bp->lines.end(), l, l + 4);
    }
    RectF clip_box{ctx.view_box.min_x, -std::numeric_limits<float>::infinity(), ctx.view_box.max_x,
                   ctx.view_box.max_y};
    if (!clip_line_segment_to_rect(from, to, clip_box)) return;

    const float tile_size = 16.0f;
    const float tile_size_recip = 1.0f / tile_size;
    // (line_segment * recip).floor().to_i32x4()
    int32_t from_tx = cvtps(floorf(from.x * tile_size_recip)), from_ty = cvtps(floorf(from.y * tile_size_recip));
    int32_t to_tx = cvtps(floorf(to.x * tile_size_recip)), to_ty = cvtps(floorf(to.y * tile_size_recip));

    V2 vector = to - from;
    bool neg_x = vector.x < 0.0f, neg_y = vector.y < 0.0f;
    int32_t step_x = neg_x ? -1 : 1, step_y = neg_y ? -1 : 1;
    float first_cross_x = (float)(from_tx + (neg_x ? 0 : 1)) * tile_size;
    float first_cross_y = (float)(from_ty + (neg_y ? 0 : 1)) * tile_size;
    float t_max_x = (first_cross_x - from.x) / vector.x;
    float t_max_y = (first_cross_y - from.y) / vector.y;
    float t_delta_x = fabsf(tile_size / vector.x);
    float t_delta_y = fabsf(tile_size / vector.y);

    V2 current_position = from;
    int32_t tx = from_tx, ty = from_ty;
    enum { NONE = 0, X = 1, Y = 2 };
    int last_step = NONE;

    for (;;) {
        int next_step;
        if (t_max_x < t_max_y)
            next_step = X;
        else if (t_max_x > t_max_y)
            next_step = Y;
        else
            next_step = step_x > 0 ? X : Y;

        float next_t = fminf(next_step == X ? t_max_x : t_max_y, 1.0f); // f32::min

        if (tx == to_tx && ty == to_ty) next_step = NONE;

        // line_segment.sample(next_t) = from + vector() * t   (line_segment.rs:214-217,231-233)
        V2 v = to - from;
        V2 next_position = v2(from.x + v.x * next_t, from.y + v.y * next_t);
        ob.add_fill(current_position, next_position, tx, ty);

        if (step_y < 0 && next_step == Y) {
            // leaves through the top boundary
            ob.add_fill(next_position, v2((float)tx * tile_size, (float)ty * tile_size), tx, ty);
        } else if (step_y > 0 && last_step == Y) {
            // entered through the top boundary
            ob.add_fill(v2((float)tx * tile_size, (float)ty * tile_size), current_position, tx, ty);
        }

        if (step_x < 0 && last_step == X) {
            ob.adjust_backdrop(tx, ty, 1); // entered through the right boundary
        } else if (step_x > 0 && next_step == X) {
            ob.adjust_backdrop(tx, ty, -1); // leaving through the right boundary
        }

        if (next_step == NONE) break;
        if (next_step == X) {
            if (tx == to_tx) break;
            t_max_x += t_delta_x;
            t_max_y += 0.0f; // t_max += vec2f(t_delta.x(), 0.0)
            tx += step_x;
        } else {
            if (ty == to_ty) break;
            t_max_x += 0.0f;
            t_max_y += t_delta_y;
            ty += step_y;
        }
        current_position = next_position;
        last_step = next_step;
    }
}

struct Cubic {
    V2 p0, p1, p2, p3; // baseline.from, ctrl.from, ctrl.to, baseline.to
};

// CubicSegment::is_flat (content/src/segment.rs:292-300), tolerance 0.25.
static inline bool cubic_is_flat(const Cubic &c) {
    // uv = 3*ctrl - baseline - baseline - baseline.reversed()
    float u0x = ((3.0f * c.p1.x - c.p0.x) - c.p0.x) - c.p3.x;
    float u0y = ((3.0f * c.p1.y - c.p0.y) - c.p0.y) - c.p3.y;
    float u1x = ((3.0f * c.p2.x - c.p3.x) - c.p3.x) - c.p0.x;
    float u1y = ((3.0f * c.p2.y - c.p3.y) - c.p3.y) - c.p0.y;
    u0x = u0x * u0x;
    u0y = u0y * u0y;
    u1x = u1x * u1x;
    u1y = u1y * u1y;
    float mx = sse_max(u0x, u1x), my = sse_max(u0y, u1y); // uv.max(uv.zwxy())
    const float tol = 0.25f;                               // tiler.rs:30
    return mx + my <= 16.0f * tol * tol;
}

static inline V2 lerp2(V2 a, V2 b, float t) { // a + t * (b - a), per lane
    return v2(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y));
}

// CubicSegment::split(0.5) (content/src/segment.rs:307-360).
static inline void cubic_split_half(const Cubic &c, Cubic &a, Cubic &b) {
    const float t = 0.5f;
    V2 p01 = lerp2(c.p0, c.p1, t), p12 = lerp2(c.p1, c.p2, t), p23 = lerp2(c.p2, c.p3, t);
    V2 p012 = lerp2(p01, p12, t), p123 = lerp2(p12, p23, t);
    V2 p0123 = lerp2(p012, p123, t);
    a = Cubic{c.p0, p01, p012, p0123};
    b = Cubic{p0123, p123, p23, c.p3};
}

// process_segment for a cubic (renderer/src/tiler.rs:166-184): recursive halving, left first.
static void process_cubic(const Cubic &c, ObjectBuilder &ob, int depth) {
    if (cubic_is_flat(c) || depth >= 40) { // depth guard (f32 halving collapses long before): the reference would overflow its stack
        process_line_segment(c.p0, c.p3, ob);
        return;
    }
    Cubic a, b;
    cubic_split_half(c, a, b);
    process_cubic(a, ob, depth + 1);
    process_cubic(b, ob, depth + 1);
}

// Tiler::generate_fills (tiler.rs:83-91) driving ContourIter (content/src/outline.rs:1014-1074);
// every contour is closed (Scene::apply_render_options -> close_all_contours).
static void generate_fills(const Outline &o, ObjectBuilder &ob) {
    for (const Contour &c : o.contours) {
        const uint32_t len = (uint32_t)c.points.size();
        uint32_t index = 1;
        for (;;) {
            if (index == len + 1) break;
            ob.bp->n_input_segments++;
            V2 point0 = c.points[index - 1];
            if (index == len) { // closing line
                index++;
                process_line_segment(point0, c.points[0], ob);
                continue;
            }
            uint32_t i1 = index++;
            V2 point1 = c.points[i1];
            if (c.flags[i1] == 0) { // point_is_endpoint
                process_line_segment(point0, point1, ob);
                continue;
            }
            uint32_t i2 = index++;
            if (i2 >= len) { // malformed contour (control point at the end): the reference would index out of bounds
                break;
            }
            V2 point2 = c.points[i2];
            if (c.flags[i2] == 0) {
                // quadratic -> Segment::to_cubic (segment.rs:171-183)
                V2 p1_2 = point1 + point1;
                const float third = 1.0f / 3.0f;
                Cubic cu{point0, (point0 + p1_2) * third, (p1_2 + point2) * third, point2};
                process_cubic(cu, ob, 0);
                continue;
            }
            uint32_t i3 = index++;
            if (i3 >= len) break;
            Cubic cu{point0, point1, point2, c.points[i3]};
            process_cubic(cu, ob, 0);
        }
    }
}

// Tiler::prepare_tiles (tiler.rs:93-163).
static void prepare_tiles(BuiltPath &bp, const BuiltPath *clip) {
    const int32_t tiles_across = bp.tile_bounds.width();
    for (size_t i = 0; i < bp.tiles.size(); i++) {
        PFOTileObjectPrimitive &draw_tile = bp.tiles[i];
        size_t column = i % (size_t)tiles_across;
        int32_t delta = (int32_t)draw_tile.backdrop;
        uint32_t draw_alpha = draw_tile.alpha_tile_id;
        int8_t draw_backdrop = (int8_t)bp.backdrops[column];
        if (clip) {
            const RectI &cr = clip->tile_bounds;
            int32_t tx = draw_tile.tile_x, ty = draw_tile.tile_y;
            if (cr.contains_point(tx, ty)) {
                const PFOTileObjectPrimitive &clip_tile =
                    clip->tiles[(size_t)(ty - cr.min_y) * cr.width() + (tx - cr.min_x)];
                if (clip_tile.alpha_tile_id != INVALID_ALPHA && draw_alpha != INVALID_ALPHA) {
                    PFOClip &cl = bp.clip_tiles[i];
                    cl.dest_tile_id = draw_tile.alpha_tile_id;
                    cl.dest_backdrop = (int32_t)draw_backdrop;
                    cl.src_tile_id = clip_tile.alpha_tile_id;
                    cl.src_backdrop = (int32_t)clip_tile.backdrop;
                    draw_backdrop = 0;
                } else if (clip_tile.alpha_tile_id != INVALID_ALPHA && draw_alpha == INVALID_ALPHA &&
                           draw_backdrop != 0) {
                    draw_alpha = clip_tile.alpha_tile_id;
                    draw_backdrop = clip_tile.backdrop;
                } else if (clip_tile.alpha_tile_id == INVALID_ALPHA && clip_tile.backdrop == 0) {
                    draw_alpha = INVALID_ALPHA;
                    draw_backdrop = 0;
                }
            } else {
                draw_alpha = INVALID_ALPHA;
                draw_backdrop = 0;
            }
        }
        draw_tile.alpha_tile_id = draw_alpha;
        draw_tile.backdrop = draw_backdrop;
        bp.backdrops[column] += delta;
    }
}

// Tiler::new + BuiltPath::new + generate_tiles for one path (tiler.rs:40-81, builder.rs:414-484).
static void build_path(const PFOScene &s, const PFOBuildOptions &opt, SceneCtx &ctx, bool is_clip,
                       uint32_t index, BuiltPath &bp, const std::vector<BuiltPath> *clips) {
    const uint32_t *range = is_clip ? &s.clip_contour_ranges[2 * index] : &s.draw_contour_ranges[2 * index];
    Outline original = load_outline(s, range[0], range[1]);
    Outline outline = apply_render_options(original, opt);

    RectF bounds;
    if (!rect_intersection(outline.bounds, ctx.effective_view_box, bounds)) bounds = RectF{0, 0, 0, 0};
    // blend modes other than SrcOver are out of scope, so has_destructive_blend_mode() is false.
    bp.tile_bounds = round_rect_out_to_tile_bounds(bounds);
    if (ctx.strip_y1 > ctx.strip_y0) { // multi-GPU strip restriction (not a reference feature)
        bp.tile_bounds.min_y = std::max(bp.tile_bounds.min_y, ctx.strip_y0);
        bp.tile_bounds.max_y = std::max(bp.tile_bounds.min_y, std::min(bp.tile_bounds.max_y, ctx.strip_y1));
    }
    bp.fill_rule = is_clip ? s.clip_fill_rules[index] : s.draw_fill_rules[index];
    uint16_t paint = is_clip ? 0 : s.draw_paints[index];
    // TilingPathInfo::to_ctrl (tiles.rs:45-61): clip paths get ctrl 0.
    bp.ctrl = is_clip ? 0 : (bp.fill_rule == PFO_FILL_RULE_EVEN_ODD ? TILE_CTRL_MASK_EVEN_ODD : TILE_CTRL_MASK_WINDING);
    const RectI &r = bp.tile_bounds;
    int32_t w = std::max(r.width(), 0), h = std::max(r.height(), 0);
    bp.backdrops.assign((size_t)w, 0);
    bp.tiles.resize((size_t)w * h);
    for (int32_t y = 0; y < h; y++)
        for (int32_t x = 0; x < w; x++)
            bp.tiles[(size_t)y * w + x] = PFOTileObjectPrimitive{
                (int16_t)(r.min_x + x), (int16_t)(r.min_y + y), INVALID_ALPHA, index, paint, bp.ctrl, 0};
    uint32_t clip_id = is_clip ? PFO_NO_CLIP : (s.draw_clip_paths ? s.draw_clip_paths[index] : PFO_NO_CLIP);
    if (clip_id != PFO_NO_CLIP) bp.clip_tiles.assign(bp.tiles.size(), PFOClip{INVALID_ALPHA, 0, INVALID_ALPHA, 0});
    if (!is_clip) {
        // BuiltDrawPath::new (builder.rs:80-94): occludes = paint opaque && blend occludes (SrcOver).
        bp.occludes = s.paint_colors[4 * (size_t)paint + 3] == 255;
    }

    ObjectBuilder ob{&bp, &ctx, index};
    generate_fills(outline, ob);
    prepare_tiles(bp, clip_id != PFO_NO_CLIP ? &(*clips)[clip_id] : nullptr);
}

bool g_keep_lines = false;

} // namespace

struct PFOBuilt {
    std::vector<BuiltPath> clip_paths, draw_paths;
    std::vector<PFOFill> fills;
    std::vector<uint32_t> fill_path_offsets;
    std::vector<PFOTileObjectPrimitive> tiles;
    std::vector<PFOClip> clips;
    std::vector<int32_t> z_buffer;
    RectI z_rect{0, 0, 0, 0};
    uint32_t alpha_tile_count = 0;
    uint64_t n_lines = 0, n_input_segments = 0, n_bbox_tiles = 0;
    double seconds = 0;       // build_paths_on_cpu + build_tile_batches (the reference's cpu_build_time)
    double seconds_paths = 0; // the per-path (parallel) part alone
};

extern "C" {

void pfo_set_keep_lines(int keep) { g_keep_lines = keep != 0; }

PFOBuilt *pfo_build(const PFOScene *scene, const PFOBuildOptions *options, int n_threads) {
    const PFOScene &s = *scene;
    PFOBuildOptions opt = *options;
    auto t0 = std::chrono::steady_clock::now();
    PFOBuilt *b = new PFOBuilt();
    SceneCtx ctx;
    ctx.view_box = RectF{s.view_box[0], s.view_box[1], s.view_box[2], s.view_box[3]};
    ctx.effective_view_box = ctx.view_box;
    if (opt.subpixel_aa_enabled) { // scene.rs:276-282: view_box * vec2f(3.0, 1.0)
        ctx.effective_view_box = RectF{ctx.view_box.min_x * 3.0f, ctx.view_box.min_y * 1.0f,
                                       ctx.view_box.max_x * 3.0f, ctx.view_box.max_y * 1.0f};
    }
    ctx.keep_lines = g_keep_lines;
    ctx.strip_y0 = opt.strip_tile_y0;
    ctx.strip_y1 = opt.strip_tile_y1;

    b->clip_paths.resize(s.n_clip_paths);
    b->draw_paths.resize(s.n_draw_paths);

    // build_paths_on_cpu (builder.rs:224-259): clip paths first, then draw paths.
    auto run = [&](bool is_clip, uint32_t count, std::vector<BuiltPath> &out) {
        if (n_threads <= 1) {
            for (uint32_t i = 0; i < count; i++) build_path(s, opt, ctx, is_clip, i, out[i], &b->clip_paths);
        } else {
            std::atomic<uint32_t> next{0};
            std::vector<std::thread> workers;
            for (int t = 0; t < n_threads; t++)
                workers.emplace_back([&]() {
                    for (;;) {
                        uint32_t i = next.fetch_add(1, std::memory_order_relaxed);
                        if (i >= count) break;
                        build_path(s, opt, ctx, is_clip, i, out[i], &b->clip_paths);
                    }
                });
            for (auto &w : workers) w.join();
        }
    };
    run(true, s.n_clip_paths, b->clip_paths);
    run(false, s.n_draw_paths, b->draw_paths);
    b->seconds_paths = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // build_tile_batches, D3D9 level (builder.rs:886-1056): one batch (solid colours only).
    RectI zr = round_rect_out_to_tile_bounds(ctx.view_box); // builder.rs:949-953 uses scene.view_box()
    b->z_rect = zr;
    b->z_buffer.assign((size_t)std::max(zr.width(), 0) * std::max(zr.height(), 0), 0);
    size_t total_tiles = 0;
    for (const BuiltPath &bp : b->draw_paths) total_tiles += bp.tiles.size();
    b->tiles.reserve(total_tiles); // address space only: pages are touched as tiles are appended, and nothing is copied
    for (uint32_t pi = 0; pi < s.n_draw_paths; pi++) {
        const BuiltPath &bp = b->draw_paths[pi];
        for (const PFOTileObjectPrimitive &tile : bp.tiles) { // builder.rs:1013-1029
            if (tile.alpha_tile_id == INVALID_ALPHA && tile.backdrop == 0) continue;
            b->tiles.push_back(tile);
            if (!bp.occludes || tile.alpha_tile_id != INVALID_ALPHA) continue;
            if (!zr.contains_point(tile.tile_x, tile.tile_y)) {
                fprintf(stderr, "pf_oracle: Z value out of bounds!\n");
                abort();
            }
            int32_t &z = b->z_buffer[(size_t)(tile.tile_y - zr.min_y) * zr.width() + (tile.tile_x - zr.min_x)];
            z = std::max(z, (int32_t)pi);
        }
        for (const PFOClip &cl : bp.clip_tiles) // builder.rs:1031-1040
            if (cl.dest_tile_id != INVALID_ALPHA && cl.src_tile_id != INVALID_ALPHA) b->clips.push_back(cl);
    }

    // cpu_build_time stops here: the reference hands each path's Vec<Fill> to the listener by move
    // (builder.rs:320-325); concatenating them below is only this oracle's output format.
    b->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // AddFillsD3D9 payloads, in path order (builder.rs:320-325).
    size_t nf = 0;
    for (const BuiltPath &bp : b->clip_paths) nf += bp.fills.size();
    for (const BuiltPath &bp : b->draw_paths) nf += bp.fills.size();
    b->fills.reserve(nf);
    b->fill_path_offsets.reserve(s.n_clip_paths + s.n_draw_paths + 1);
    auto gather = [&](std::vector<BuiltPath> &paths) {
        for (BuiltPath &bp : paths) {
            b->fill_path_offsets.push_back((uint32_t)b->fills.size());
            b->fills.insert(b->fills.end(), bp.fills.begin(), bp.fills.end());
            b->n_lines += bp.n_lines;
            b->n_input_segments += bp.n_input_segments;
            b->n_bbox_tiles += bp.tiles.size();
        }
    };
    gather(b->clip_paths);
    gather(b->draw_paths);
    b->fill_path_offsets.push_back((uint32_t)b->fills.size());
    b->alpha_tile_count = ctx.next_alpha_tile.load();
    return b;
}

void pfo_built_destroy(PFOBuilt *b) { delete b; }
size_t pfo_fill_count(const PFOBuilt *b) { return b->fills.size(); }
const PFOFill *pfo_fills(const PFOBuilt *b) { return b->fills.data(); }
const uint32_t *pfo_fill_path_offsets(const PFOBuilt *b) { return b->fill_path_offsets.data(); }
size_t pfo_tile_count(const PFOBuilt *b) { return b->tiles.size(); }
const PFOTileObjectPrimitive *pfo_tiles(const PFOBuilt *b) { return b->tiles.data(); }
size_t pfo_clip_count(const PFOBuilt *b) { return b->clips.size(); }
const PFOClip *pfo_clips(const PFOBuilt *b) { return b->clips.data(); }
const int32_t *pfo_z_buffer(const PFOBuilt *b, int32_t rect_out[4]) {
    if (rect_out) {
        rect_out[0] = b->z_rect.min_x;
        rect_out[1] = b->z_rect.min_y;
        rect_out[2] = b->z_rect.max_x;
        rect_out[3] = b->z_rect.max_y;
    }
    return b->z_buffer.data();
}
uint32_t pfo_alpha_tile_count(const PFOBuilt *b) { return b->alpha_tile_count; }
uint64_t pfo_line_segment_count(const PFOBuilt *b) { return b->n_lines; }
uint64_t pfo_input_segment_count(const PFOBuilt *b) { return b->n_input_segments; }
uint64_t pfo_bbox_tile_count(const PFOBuilt *b) { return b->n_bbox_tiles; }
double pfo_build_seconds(const PFOBuilt *b) { return b->seconds; }
double pfo_build_seconds_paths(const PFOBuilt *b) { return b->seconds_paths; }

size_t pfo_path_lines(const PFOBuilt *b, uint32_t path, float *out, size_t cap) {
    const BuiltPath &bp = path < b->clip_paths.size() ? b->clip_paths[path] : b->draw_paths[path - b->clip_paths.size()];
    size_t n = bp.lines.size() / 4;
    if (out) memcpy(out, bp.lines.data(), std::min(n, cap) * 4 * sizeof(float));
    return n;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// Coverage + composite (the "CPU evaluation of the reference's fill math").
// ---------------------------------------------------------------------------------------------

namespace {

// texture(areaLUT, uv) with LINEAR filtering and CLAMP_TO_EDGE on a 256x256 RGBA8 texture
// (renderer/src/gpu/renderer.rs:207-214; gl/src/lib.rs:375,674-705), exact float bilinear.
static inline void sample_lut(const uint8_t *lut, float u, float v, float out[4]) {
    float x = u * 256.0f - 0.5f, y = v * 256.0f - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    auto cl = [](int i) { return i < 0 ? 0 : (i > 255 ? 255 : i); };
    x0 = cl(x0), x1 = cl(x1), y0 = cl(y0), y1 = cl(y1);
    const uint8_t *t00 = lut + 4 * (y0 * 256 + x0), *t10 = lut + 4 * (y0 * 256 + x1);
    const uint8_t *t01 = lut + 4 * (y1 * 256 + x0), *t11 = lut + 4 * (y1 * 256 + x1);
    for (int k = 0; k < 4; k++) {
        float a = (float)t00[k] * (1.0f / 255.0f), b = (float)t10[k] * (1.0f / 255.0f);
        float c = (float)t01[k] * (1.0f / 255.0f), d = (float)t11[k] * (1.0f / 255.0f);
        float top = a + (b - a) * ax, bot = c + (d - c) * ax;
        out[k] = top + (bot - top) * ay;
    }
}

static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; } // GLSL mix

// computeCoverage (shaders/fill_area.inc.glsl:11-27). from/to are relative to the centre of the
// first pixel of a 4-row strip; returns the coverage of the 4 rows.
static inline void compute_coverage(V2 from, V2 to, const uint8_t *lut, float out[4]) {
    V2 left = from.x < to.x ? from : to, right = from.x < to.x ? to : from;
    float wx = clampf(from.x, -0.5f, 0.5f), wy = clampf(to.x, -0.5f, 0.5f);
    float offset = mixf(wx, wy, 0.5f) - left.x;
    float t = offset / (right.x - left.x);
    float y = mixf(left.y, right.y, t);
    float d = (right.y - left.y) / (right.x - left.x);
    float dX = wx - wy;
    float tex[4];
    sample_lut(lut, (y + 8.0f) / 16.0f, fabsf(d * dX) / 16.0f, tex);
    for (int k = 0; k < 4; k++) out[k] = tex[k] * dX;
}

static inline float half_round(float f) {
    // f32 -> f16 -> f32, round-to-nearest-even (half::f16::from_f32, gpu/renderer.rs:714-762).
    _Float16 h = (_Float16)f;
    return (float)h;
}

// One fill's contribution to its alpha tile's mask (16 x 16 f32): D3D9 draws one instanced quad per fill
// with additive blending (d3d9/renderer.rs:237-263); D3D11: accumulateCoverageForFillList
// (fill_compute.inc.glsl:11-25).
static void add_fill_to_mask(const PFOFill &f, const uint8_t *lut, float *mask) {
    // lineSegment = vec4(packed) / 256.0
    V2 from = v2((float)f.from_x / 256.0f, (float)f.from_y / 256.0f);
    V2 to = v2((float)f.to_x / 256.0f, (float)f.to_y / 256.0f);
    for (int strip = 0; strip < 4; strip++) {
        for (int x = 0; x < 16; x++) {
            // tileFragCoord = vec2(tileSubCoord) + 0.5 with tileSubCoord = (x, 4*strip)
            V2 c = v2((float)x + 0.5f, (float)(4 * strip) + 0.5f);
            float cov[4];
            compute_coverage(from - c, to - c, lut, cov);
            for (int k = 0; k < 4; k++) mask[(4 * strip + k) * 16 + x] += cov[k];
        }
    }
}

} // namespace

extern "C" {

void pfo_alpha_masks(const PFOBuilt *b, const uint8_t *lut, float *out) {
    size_t n = (size_t)b->alpha_tile_count * 256;
    for (size_t i = 0; i < n; i++) out[i] = 0.0f;
    // Fills are accumulated in emission order.
    for (const PFOFill &f : b->fills) add_fill_to_mask(f, lut, out + (size_t)f.link * 256);
}

// Composites the pixels [x0, x0 + width) x [y0, y0 + height) of the frame (frame_w x frame_h pixels). Only the
// masks of alpha tiles that reach the crop are evaluated, so a crop of a frame too large to hold every mask
// (1M paths at 16384^2) costs what its own tiles cost.
void pfo_render_crop(const PFOBuilt *b, const PFOScene *scene, const uint8_t *lut, const float background[4],
                     uint32_t frame_w, uint32_t frame_h, uint32_t x0, uint32_t y0, uint32_t width, uint32_t height,
                     uint8_t *out_rgba, float *out_f32) {
    const RectI zr = b->z_rect;
    const int64_t cx0 = x0, cy0 = y0, cx1 = std::min<int64_t>((int64_t)x0 + width, frame_w),
                  cy1 = std::min<int64_t>((int64_t)y0 + height, frame_h);
    auto tile_drawn = [&](const PFOTileObjectPrimitive &tile) {
        if (!zr.contains_point(tile.tile_x, tile.tile_y)) return false;
        const int64_t tx = (int64_t)tile.tile_x * 16, ty = (int64_t)tile.tile_y * 16;
        if (tx + 16 <= cx0 || tx >= cx1 || ty + 16 <= cy0 || ty >= cy1) return false;
        // culled when path_id < z (d3d9/tile.vs.glsl:52-56 / sort.cs.glsl:74)
        int32_t z = b->z_buffer[(size_t)(tile.tile_y - zr.min_y) * zr.width() + (tile.tile_x - zr.min_x)];
        return (int32_t)tile.path_id >= z;
    };
    // Which alpha tiles are needed: those of drawn tiles, and for every Clip record whose destination is
    // needed, its source.
    const uint32_t NONE = 0xffffffffu;
    std::vector<uint32_t> slot(b->alpha_tile_count, NONE);
    uint32_t n_slots = 0;
    for (const PFOTileObjectPrimitive &tile : b->tiles)
        if (tile.alpha_tile_id != INVALID_ALPHA && tile_drawn(tile) && slot[tile.alpha_tile_id] == NONE)
            slot[tile.alpha_tile_id] = n_slots++;
    for (const PFOClip &cl : b->clips)
        if (slot[cl.dest_tile_id] != NONE && slot[cl.src_tile_id] == NONE) slot[cl.src_tile_id] = n_slots++;
    std::vector<float> masks((size_t)n_slots * 256, 0.0f);
    for (const PFOFill &f : b->fills) // emission order
        if (slot[f.link] != NONE) add_fill_to_mask(f, lut, masks.data() + (size_t)slot[f.link] * 256);
    // D3D9 clip combine (shaders/d3d9/tile_clip_combine.fs.glsl:28-31):
    // dest = min(abs(dest + dest_backdrop), abs(src + src_backdrop)); the draw tile's backdrop was
    // zeroed by prepare_tiles.
    for (const PFOClip &cl : b->clips) {
        if (slot[cl.dest_tile_id] == NONE) continue;
        float *dst = masks.data() + (size_t)slot[cl.dest_tile_id] * 256;
        const float *src = masks.data() + (size_t)slot[cl.src_tile_id] * 256;
        for (int i = 0; i < 256; i++)
            dst[i] = fminf(fabsf(dst[i] + (float)cl.dest_backdrop), fabsf(src[i] + (float)cl.src_backdrop));
    }

    std::vector<float> dest((size_t)width * height * 4);
    for (size_t i = 0; i < (size_t)width * height; i++)
        for (int k = 0; k < 4; k++) dest[4 * i + k] = background[k]; // LOAD_ACTION_CLEAR

    // Paint table: ColorU::to_f32 (color/src/lib.rs:70-73) then f16 (gpu/renderer.rs:726-729).
    std::vector<float> paint((size_t)scene->n_paints * 4);
    for (size_t i = 0; i < paint.size(); i++) paint[i] = half_round((float)scene->paint_colors[i] * (1.0f / 255.0f));

    // Tiles are in ascending draw order = painter's order (sort.cs.glsl:60-95).
    for (const PFOTileObjectPrimitive &tile : b->tiles) {
        if (!tile_drawn(tile)) continue;
        const float *base = &paint[4 * (size_t)tile.color];
        const float *mask = tile.alpha_tile_id != INVALID_ALPHA ? masks.data() + (size_t)slot[tile.alpha_tile_id] * 256 : nullptr;
        for (int py = 0; py < 16; py++) {
            int64_t y = (int64_t)tile.tile_y * 16 + py;
            if (y < cy0 || y >= cy1) continue;
            for (int px = 0; px < 16; px++) {
                int64_t x = (int64_t)tile.tile_x * 16 + px;
                if (x < cx0 || x >= cx1) continue;
                // sampleMask (tile_fragment.inc.glsl:539-556): coverage = texel + backdrop, then rule.
                float coverage = (mask ? mask[py * 16 + px] : 0.0f) + (float)tile.backdrop;
                float mask_alpha;
                if (tile.ctrl & TILE_CTRL_MASK_WINDING) {
                    coverage = fabsf(coverage);
                } else if (tile.ctrl & TILE_CTRL_MASK_EVEN_ODD) {
                    float m = coverage - 2.0f * floorf(coverage / 2.0f); // GLSL mod(x, 2.0)
                    coverage = 1.0f - fabsf(1.0f - m);
                } else {
                    coverage = 1.0f; // maskCtrl == 0: no mask
                }
                mask_alpha = fminf(1.0f, coverage);
                // calculateColor (tile_fragment.inc.glsl:560-614), solid colour, SrcOver.
                float a = base[3] * mask_alpha;
                float src[4] = {base[0] * a, base[1] * a, base[2] * a, a};
                float *d = &dest[4 * ((size_t)(y - cy0) * width + (size_t)(x - cx0))];
                for (int k = 0; k < 4; k++) d[k] = d[k] * (1.0f - a) + src[k]; // tile.cs.glsl:155
            }
        }
    }
    for (size_t i = 0; i < dest.size(); i++) {
        if (out_f32) out_f32[i] = dest[i];
        // imageStore to rgba8: clamp to [0,1], * 255, round to nearest.
        float v = fminf(fmaxf(dest[i], 0.0f), 1.0f) * 255.0f;
        out_rgba[i] = (uint8_t)lrintf(v);
    }
}

void pfo_render(const PFOBuilt *b, const PFOScene *scene, const uint8_t *lut, const float background[4],
                uint32_t width, uint32_t height, uint8_t *out_rgba, float *out_f32) {
    pfo_render_crop(b, scene, lut, background, width, height, 0, 0, width, height, out_rgba, out_f32);
}

} // extern "C"
