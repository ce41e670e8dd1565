// pathfinder_b200/csrc/stroke.cpp — stroke-to-fill on the host: a C++ restatement of
// OutlineStrokeToFill (content/src/stroke.rs:88-448) in the reference's f32 arithmetic, one rounding per
// operation (compiled with -ffp-contract=off). Miter, bevel and round joins; butt, square and round caps (round
// shapes through Contour::push_arc_from_unit_chord, content/src/outline.rs:632-680: at most four cubic arcs of at
// most a quarter circle each, content/src/segment.rs:94-122).
//
// Geometry helpers follow geometry/src/vector.rs:118-126 (length, normalize = v * (1 / length)),
// geometry/src/line_segment.rs:219-247 (intersection_t, sample, offset), geometry/src/transform2d.rs:60-72,
// 123-130 (adjugate, det, inverse, matrix * vector) and content/src/segment.rs:171-233,307-380 (to_cubic,
// reversed, sample, de Casteljau split).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pf_cuda.h"
#include "outline.h"

namespace pf {
void set_last_error(const std::string &msg);
}

namespace {

struct V2 {
    float x, y;
};
inline V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
inline V2 operator*(V2 a, float s) { return V2{a.x * s, a.y * s}; }
inline V2 operator*(V2 a, V2 b) { return V2{a.x * b.x, a.y * b.y}; }
inline bool operator==(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
inline float square_length(V2 a) { return a.x * a.x + a.y * a.y; }
inline V2 normalize(V2 a) { return a * (1.0f / std::sqrt(square_length(a))); }
inline V2 lerp(V2 a, V2 b, float t) { return a + (b - a) * t; }

constexpr float TOLERANCE = 0.01f; // stroke.rs:22
constexpr float EPSILON = 0.001f;  // geometry/src/util.rs:15

struct Line {
    V2 from, to;
    V2 vector() const { return to - from; }
    V2 sample(float t) const { return from + vector() * t; }
    float square_length() const { return ::square_length(vector()); }
};

// LineSegment2F::offset (line_segment.rs:241-247)
Line offset_line(Line l, float distance) {
    const V2 v = l.vector();
    if (v.x == 0.0f && v.y == 0.0f) return l;
    const V2 o = normalize(V2{v.y, v.x}) * V2{-distance, distance};
    return Line{l.from + o, l.to + o};
}

// LineSegment2F::intersection_t (line_segment.rs:219-228)
bool intersection_t(Line a, Line b, float &t) {
    const V2 p0p1 = a.vector(), ov = b.vector();
    const float m0 = ov.x, m1 = ov.y, m2 = -p0p1.x, m3 = -p0p1.y; // Matrix2x2F(m11, m21, m12, m22)
    const float det = m0 * m3 - m2 * m1;
    if (std::fabs(det) < 0.0001f) return false;
    const float inv_det = 1.0f / det;
    // inverse = splat(1 / det) * adjugate = (m3, -m1, -m2, m0) / det; (inverse * r).y = inv[1] * r.x + inv[3] * r.y
    const V2 r = a.from - b.from;
    t = (inv_det * (-m1)) * r.x + (inv_det * m0) * r.y;
    return true;
}

enum Kind { LINE = 0, QUADRATIC = 1, CUBIC = 2 };
struct Segment {
    Kind kind;
    Line baseline; // from, to
    Line ctrl;     // control points (quadratic: ctrl.from only)
};

// Segment::to_cubic (segment.rs:171-183)
Segment to_cubic(const Segment &s) {
    if (s.kind == CUBIC) return s;
    Segment c = s;
    const V2 p1_2 = s.ctrl.from + s.ctrl.from;
    c.ctrl = Line{(s.baseline.from + p1_2) * (1.0f / 3.0f), (p1_2 + s.baseline.to) * (1.0f / 3.0f)};
    c.kind = CUBIC;
    return c;
}

// CubicSegment::split (segment.rs:307-360)
void split_cubic(const Segment &s, float t, Segment &before, Segment &after) {
    before.kind = after.kind = CUBIC;
    if (t <= 0.0f) {
        before.baseline = Line{s.baseline.from, s.baseline.from};
        before.ctrl = before.baseline;
        after.baseline = s.baseline, after.ctrl = s.ctrl;
        return;
    }
    if (t >= 1.0f) {
        before.baseline = s.baseline, before.ctrl = s.ctrl;
        after.baseline = Line{s.baseline.to, s.baseline.to};
        after.ctrl = after.baseline;
        return;
    }
    const V2 p0 = s.baseline.from, p1 = s.ctrl.from, p2 = s.ctrl.to, p3 = s.baseline.to;
    const V2 p01 = lerp(p0, p1, t), p12 = lerp(p1, p2, t), p23 = lerp(p2, p3, t);
    const V2 p012 = lerp(p01, p12, t), p123 = lerp(p12, p23, t);
    const V2 p0123 = lerp(p012, p123, t);
    before.baseline = Line{p0, p0123}, before.ctrl = Line{p01, p012};
    after.baseline = Line{p0123, p3}, after.ctrl = Line{p123, p23};
}

// Segment::split (segment.rs:211-219)
void split(const Segment &s, float t, Segment &before, Segment &after) {
    if (s.kind == LINE) {
        const V2 mid = s.baseline.from + (s.baseline.to - s.baseline.from) * t;
        before = Segment{LINE, Line{s.baseline.from, mid}, Line{}};
        after = Segment{LINE, Line{mid, s.baseline.to}, Line{}};
        return;
    }
    split_cubic(to_cubic(s), t, before, after);
}

// Segment::sample (segment.rs:226-233, 378-380)
V2 sample(const Segment &s, float t) {
    if (s.kind == LINE) return s.baseline.sample(t);
    Segment before, after;
    split_cubic(to_cubic(s), t, before, after);
    return before.baseline.to;
}

// Segment::reversed (segment.rs:186-197)
Segment reversed(const Segment &s) {
    Segment r = s;
    r.baseline = Line{s.baseline.to, s.baseline.from};
    if (s.kind == CUBIC) r.ctrl = Line{s.ctrl.to, s.ctrl.from};
    return r;
}

// Matrix2x2F (m11, m21, m12, m22) and Transform2F, with the SIMD formulas of geometry/src/transform2d.rs:45-47,
// 115-130,301-317 written out lane by lane.
struct Mat {
    float m[4];
};
struct Xf {
    Mat matrix;
    V2 vector;
};
inline Mat mat_mul(const Mat &a, const Mat &b) {
    return Mat{{a.m[0] * b.m[0] + a.m[2] * b.m[1], a.m[1] * b.m[0] + a.m[3] * b.m[1], a.m[0] * b.m[2] + a.m[2] * b.m[3],
                a.m[1] * b.m[2] + a.m[3] * b.m[3]}};
}
inline V2 mat_apply(const Mat &a, V2 v) { return V2{a.m[0] * v.x + a.m[2] * v.y, a.m[1] * v.x + a.m[3] * v.y}; }
inline V2 xf_apply(const Xf &t, V2 v) { return mat_apply(t.matrix, v) + t.vector; }
inline Xf xf_mul(const Xf &a, const Xf &b) { return Xf{mat_mul(a.matrix, b.matrix), xf_apply(a, b.vector)}; }
// Transform2F::from_scale(scale).translate(v) = from_translation(v) * from_scale(scale)
inline Xf scale_then_translate(float scale, V2 v) {
    const Xf translation{Mat{{1.0f, 0.0f, 0.0f, 1.0f}}, v}, scaling{Mat{{scale, 0.0f, 0.0f, scale}}, V2{0.0f, 0.0f}};
    return xf_mul(translation, scaling);
}

// UnitVector (geometry/src/unit_vector.rs:25-46)
inline V2 rotate_by(V2 a, V2 b) { return V2{a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y}; }
inline V2 rev_rotate_by(V2 a, V2 b) { return V2{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }
inline V2 halve_angle(V2 a) {
    const float px = 0.5f * (1.0f + a.x), py = 0.5f * (1.0f + -a.x);
    return V2{std::sqrt(px > 0.0f ? px : 0.0f), std::sqrt(py > 0.0f ? py : 0.0f)};
}

struct Contour {
    std::vector<V2> points;
    std::vector<uint8_t> flags;
    size_t len() const { return points.size(); }
    void push_point(V2 p, uint8_t f) {
        points.push_back(p);
        flags.push_back(f);
    }
    void push_endpoint(V2 p) { push_point(p, 0); }
    // Contour::push_segment (outline.rs:576-600)
    void push_segment(const Segment &s) {
        push_point(s.baseline.from, 0);
        if (s.kind != LINE) {
            push_point(s.ctrl.from, PF_POINT_FLAGS_CONTROL_POINT_0);
            if (s.kind != QUADRATIC) push_point(s.ctrl.to, PF_POINT_FLAGS_CONTROL_POINT_1);
        }
        push_point(s.baseline.to, 0);
    }
    // Contour::push_arc_from_unit_chord (outline.rs:632-680), ArcDirection::CW.
    void push_arc_from_unit_chord(const Xf &transform, V2 chord_from, V2 chord_to);
    // stroke.rs:383-392
    bool might_need_join(uint32_t join) const { return len() >= 2 && join != PF_LINE_JOIN_BEVEL; }
    // stroke.rs:394-431
    void add_join(float distance, uint32_t join, float miter_limit, V2 join_point, Line next_tangent) {
        const Line prev_tangent{points[len() - 2], points[len() - 1]};
        if (prev_tangent.square_length() < EPSILON || next_tangent.square_length() < EPSILON) return;
        if (join == PF_LINE_JOIN_ROUND) {
            const Xf transform = scale_then_translate(std::fabs(distance), join_point);
            push_arc_from_unit_chord(transform, normalize(prev_tangent.to - join_point), normalize(next_tangent.to - join_point));
            return;
        }
        if (join != PF_LINE_JOIN_MITER) return;
        float t;
        if (!intersection_t(prev_tangent, next_tangent, t)) return;
        if (t < -EPSILON) return;
        const V2 miter_endpoint = prev_tangent.sample(t);
        const float threshold = miter_limit * distance;
        if (::square_length(miter_endpoint - join_point) > threshold * threshold) return;
        push_endpoint(miter_endpoint);
    }
};

// Segment::arc_from_cos (segment.rs:94-111): a unit arc centred on the +x axis, from below it to above it.
Segment arc_from_cos(float c) {
    if (c >= 1.0f - EPSILON) return Segment{LINE, Line{V2{1.0f, 0.0f}, V2{1.0f, 0.0f}}, Line{}};
    const float p0x = std::sqrt((1.0f + c) * 0.5f), p0y = std::sqrt((1.0f + -c) * 0.5f);
    const float p1x = 4.0f - p0x, p1y = (1.0f - p0x) * (3.0f - p0x) / p0y;
    const float third = 1.0f / 3.0f;
    return Segment{CUBIC, Line{V2{p0x, -p0y}, V2{p0x, p0y}}, Line{V2{p1x * third, -p1y * third}, V2{p1x * third, p1y * third}}};
}
// Segment::quarter_circle_arc (segment.rs:116-122)
Segment quarter_circle_arc() {
    const float sqrt2 = 1.41421356237309504880168872420969808f;
    const V2 p0{sqrt2 * 0.5f, sqrt2 * 0.5f};
    const V2 p1{-sqrt2 / 6.0f + 4.0f / 3.0f, 7.0f * sqrt2 / 6.0f - 4.0f / 3.0f};
    return Segment{CUBIC, Line{V2{p0.x, -p0.y}, p0}, Line{V2{p1.x, -p1.y}, p1}};
}

void Contour::push_arc_from_unit_chord(const Xf &transform, V2 chord_from, V2 chord_to) {
    V2 vector = chord_from;
    const V2 end_vector = chord_to;
    for (int segment_index = 0; segment_index < 4; segment_index++) {
        V2 sweep = rev_rotate_by(end_vector, vector);
        const bool last = sweep.x >= -EPSILON && sweep.y >= -EPSILON;
        Segment segment;
        if (!last) {
            sweep = V2{0.0f, 1.0f};
            segment = quarter_circle_arc();
        } else {
            segment = arc_from_cos(sweep.x);
        }
        const V2 r = rotate_by(halve_angle(sweep), vector);
        const Xf rotation{Mat{{r.x, r.y, -r.y, r.x}}, V2{0.0f, 0.0f}}; // Matrix2x2F::from_rotation_vector
        const Xf identity{Mat{{1.0f, 0.0f, 0.0f, 1.0f}}, V2{0.0f, 0.0f}};
        const Xf t = xf_mul(xf_mul(transform, identity), rotation); // transform * direction_transform * rotation
        segment.baseline = Line{xf_apply(t, segment.baseline.from), xf_apply(t, segment.baseline.to)};
        if (segment.kind != LINE) segment.ctrl = Line{xf_apply(t, segment.ctrl.from), xf_apply(t, segment.ctrl.to)};
        push_segment(segment);
        if (last) break;
        vector = rotate_by(vector, sweep);
    }
}

struct Stroker {
    uint32_t join;
    float miter_limit;
    // The reference's recursion (stroke.rs:243-261) only ends when the error is within tolerance or the piece is
    // shorter than the tolerance; on geometry where neither happens soon (cusps a million pixels long) it produces
    // pieces without bound. The C ABI gives up instead once one call has emitted this many pieces.
    static constexpr size_t MAX_PIECES = 1u << 20;
    mutable size_t pieces = 0;
    struct TooManyPieces {};

    // Offset::offset_once (stroke.rs:286-348)
    static V2 control_point(Line s0, Line s1) {
        float t;
        return intersection_t(s0, s1, t) ? s0.sample(t) : lerp(s0.to, s1.from, 0.5f);
    }
    static Segment offset_once(const Segment &s, float d) {
        if (s.kind == LINE) return Segment{LINE, offset_line(s.baseline, d), Line{}};
        if (s.kind == QUADRATIC) {
            const Line s0 = offset_line(Line{s.baseline.from, s.ctrl.from}, d), s1 = offset_line(Line{s.ctrl.from, s.baseline.to}, d);
            return Segment{QUADRATIC, Line{s0.from, s1.to}, Line{control_point(s0, s1), V2{0, 0}}};
        }
        if (s.baseline.from == s.ctrl.from) {
            const Line s0 = offset_line(Line{s.baseline.from, s.ctrl.to}, d), s1 = offset_line(Line{s.ctrl.to, s.baseline.to}, d);
            return Segment{CUBIC, Line{s0.from, s1.to}, Line{s0.from, control_point(s0, s1)}};
        }
        if (s.ctrl.to == s.baseline.to) {
            const Line s0 = offset_line(Line{s.baseline.from, s.ctrl.from}, d), s1 = offset_line(Line{s.ctrl.from, s.baseline.to}, d);
            return Segment{CUBIC, Line{s0.from, s1.to}, Line{control_point(s0, s1), s1.to}};
        }
        const Line s0 = offset_line(Line{s.baseline.from, s.ctrl.from}, d), s1 = offset_line(Line{s.ctrl.from, s.ctrl.to}, d),
                   s2 = offset_line(Line{s.ctrl.to, s.baseline.to}, d);
        float t0, t1;
        V2 c0, c1;
        if (intersection_t(s0, s1, t0) && intersection_t(s1, s2, t1)) {
            c0 = s0.sample(t0), c1 = s1.sample(t1);
        } else {
            c0 = lerp(s0.to, s1.from, 0.5f), c1 = lerp(s1.to, s2.from, 0.5f);
        }
        return Segment{CUBIC, Line{s0.from, s2.to}, Line{c0, c1}};
    }

    // Offset::error_is_within_tolerance (stroke.rs:350-377)
    static bool error_is_within_tolerance(const Segment &s, const Segment &other, float distance) {
        float min = std::fabs(distance) - TOLERANCE, max = std::fabs(distance) + TOLERANCE;
        min = min <= 0.0f ? 0.0f : min * min;
        max = max <= 0.0f ? 0.0f : max * max;
        for (uint32_t i = 0; i <= 16; i++) {
            const float t = (float)i / 16.0f;
            const float sq = ::square_length(sample(s, t) - sample(other, t));
            if (sq < min || sq > max) return false;
        }
        return true;
    }

    // Offset::add_to_contour (stroke.rs:263-284)
    void add_to_contour(const Segment &s, float distance, uint32_t seg_join, V2 join_point, Contour &out) const {
        if (out.might_need_join(seg_join)) {
            const V2 p3 = s.baseline.from, p4 = s.kind == LINE ? s.baseline.to : s.ctrl.from;
            out.add_join(distance, seg_join, miter_limit, join_point, Line{p4, p3});
        }
        out.push_segment(s);
    }

    // Offset::offset (stroke.rs:243-261)
    void offset(const Segment &s, float distance, uint32_t seg_join, Contour &out, int depth = 0) const {
        const V2 join_point = s.baseline.from;
        if (++pieces > MAX_PIECES) throw TooManyPieces{};
        if (s.baseline.square_length() < TOLERANCE * TOLERANCE || depth > 64) {
            add_to_contour(s, distance, seg_join, join_point, out);
            return;
        }
        const Segment candidate = offset_once(s, distance);
        if (error_is_within_tolerance(s, candidate, distance)) {
            add_to_contour(candidate, distance, seg_join, join_point, out);
            return;
        }
        Segment before, after;
        split(s, 0.5f, before, after);
        offset(before, distance, seg_join, out, depth + 1);
        offset(after, distance, seg_join, out, depth + 1);
    }
};

// ContourIter (outline.rs:1019-1062) with the close segment of closed contours.
std::vector<Segment> contour_segments(const V2 *pts, const uint8_t *flags, uint32_t n, bool closed) {
    std::vector<Segment> out;
    uint32_t index = 1;
    while (true) {
        if ((index == n && !closed) || index == n + 1) break;
        const V2 p0 = pts[index - 1];
        if (index == n) {
            out.push_back(Segment{LINE, Line{p0, pts[0]}, Line{}});
            index++;
            continue;
        }
        const uint32_t i1 = index++;
        if (flags[i1] == 0) {
            out.push_back(Segment{LINE, Line{p0, pts[i1]}, Line{}});
            continue;
        }
        const uint32_t i2 = index++;
        if (i2 >= n) break; // malformed: a control point at the end
        if (flags[i2] == 0) {
            out.push_back(Segment{QUADRATIC, Line{p0, pts[i2]}, Line{pts[i1], V2{0, 0}}});
            continue;
        }
        const uint32_t i3 = index++;
        if (i3 >= n) break;
        out.push_back(Segment{CUBIC, Line{p0, pts[i3]}, Line{pts[i1], pts[i2]}});
    }
    return out;
}

// OutlineStrokeToFill::add_cap (stroke.rs:151-199).
void add_cap(Contour &c, uint32_t cap, float width) {
    if (cap == PF_LINE_CAP_BUTT || c.len() < 2) return;
    const V2 p1 = c.points[c.len() - 1];
    V2 p0;
    size_t i = c.len() - 2;
    for (;;) {
        p0 = c.points[i];
        if (square_length(p1 - p0) > EPSILON) break;
        if (i == 0) return;
        i--;
    }
    const V2 gradient = normalize(p1 - p0);
    if (cap == PF_LINE_CAP_ROUND) {
        const V2 offset = V2{gradient.y, gradient.x} * V2{-1.0f, 1.0f};
        const Xf transform = scale_then_translate(width * 0.5f, p1 + offset * (width * 0.5f));
        c.push_arc_from_unit_chord(transform, V2{-offset.x, -offset.y}, offset);
        return;
    }
    const V2 offset = gradient * (width * 0.5f);
    const V2 p2 = p1 + offset;
    const V2 p3 = p2 + V2{gradient.y, gradient.x} * V2{-width, width};
    const V2 p4 = p3 - offset;
    c.push_endpoint(p2);
    c.push_endpoint(p3);
    c.push_endpoint(p4);
}

} // namespace

extern "C" {

// OutlineStrokeToFill::{new, offset, into_outline} (stroke.rs:88-131).
PFOutlineRef PFOutlineStrokeToFill(const PFVector2F *points, const uint8_t *point_flags, const uint32_t *contour_offsets,
                                   const uint8_t *contour_closed, uint32_t contour_count, const PFStrokeStyle *style) {
    if (!style || (contour_count && (!points || !point_flags || !contour_offsets || !contour_closed))) {
        pf::set_last_error("PFOutlineStrokeToFill: null argument");
        return nullptr;
    }
    if (style->line_cap > PF_LINE_CAP_ROUND || style->line_join > PF_LINE_JOIN_ROUND) {
        pf::set_last_error("PFOutlineStrokeToFill: unknown line cap or line join");
        return nullptr;
    }
    // Coordinates whose squares overflow f32 (or that are not numbers at all) can never meet the tolerance test: the
    // reference would recurse until its stack ran out. Refuse them here.
    const size_t point_count = contour_count ? contour_offsets[contour_count] : 0;
    for (size_t i = 0; i < point_count; i++) {
        if (!(std::fabs(points[i].x) < 1e18f) || !(std::fabs(points[i].y) < 1e18f)) {
            pf::set_last_error("PFOutlineStrokeToFill: coordinates must be finite and below 1e18 in magnitude");
            return nullptr;
        }
    }
    if (!(std::fabs(style->line_width) < 1e18f) || !(std::fabs(style->miter_limit) < 1e18f)) {
        pf::set_last_error("PFOutlineStrokeToFill: line width and miter limit must be finite");
        return nullptr;
    }
    const float radius = style->line_width * 0.5f;
    const Stroker stroker{style->line_join, style->miter_limit};
    PFOutline *out = new PFOutline;
    try {
    auto push_contour = [&](Contour &c, bool closed, V2 input_first_point) { // push_stroked_contour (stroke.rs:133-149)
        if (closed && c.might_need_join(style->line_join)) {
            const V2 p1 = c.points[1], p0 = c.points[0];
            c.add_join(radius, style->line_join, style->miter_limit, input_first_point, Line{p1, p0});
        }
        for (size_t i = 0; i < c.len(); i++) {
            out->points.push_back(PFVector2F{c.points[i].x, c.points[i].y});
            out->flags.push_back(c.flags[i]);
        }
        out->contour_offsets.push_back((uint32_t)out->points.size());
        out->closed.push_back(1); // stroker.output.closed = true (stroke.rs:148)
    };
    for (uint32_t ci = 0; ci < contour_count; ci++) {
        const uint32_t p0 = contour_offsets[ci], n = contour_offsets[ci + 1] - p0;
        if (n == 0) continue;
        const bool closed = contour_closed[ci] != 0;
        const V2 *pts = reinterpret_cast<const V2 *>(points) + p0;
        const std::vector<Segment> segments = contour_segments(pts, point_flags + p0, n, closed);
        Contour c;
        // offset_forward / offset_backward (stroke.rs:216-240): the first segment of each pass gets a bevel join
        for (size_t i = 0; i < segments.size(); i++)
            stroker.offset(segments[i], -radius, i == 0 ? (uint32_t)PF_LINE_JOIN_BEVEL : style->line_join, c);
        if (closed) {
            push_contour(c, true, pts[0]);
            c = Contour{};
        } else {
            add_cap(c, style->line_cap, style->line_width);
        }
        for (size_t i = 0; i < segments.size(); i++)
            stroker.offset(reversed(segments[segments.size() - 1 - i]), -radius,
                           i == 0 ? (uint32_t)PF_LINE_JOIN_BEVEL : style->line_join, c);
        if (!closed) add_cap(c, style->line_cap, style->line_width);
        push_contour(c, closed, pts[0]);
    }
    } catch (const Stroker::TooManyPieces &) {
        pf::set_last_error("PFOutlineStrokeToFill: the offset curves did not converge (more than 1M pieces)");
        delete out;
        return nullptr;
    } catch (const std::exception &e) {
        pf::set_last_error(std::string("PFOutlineStrokeToFill: ") + e.what());
        delete out;
        return nullptr;
    }
    return out;
}

uint32_t PFOutlineGetContourCount(PFOutlineRef outline) { return outline ? (uint32_t)outline->contour_offsets.size() - 1 : 0; }
size_t PFOutlineGetPointCount(PFOutlineRef outline) { return outline ? outline->points.size() : 0; }

void PFOutlineCopy(PFOutlineRef outline, PFVector2F *points, uint8_t *point_flags, uint32_t *contour_offsets) {
    if (!outline) return;
    if (points && !outline->points.empty()) memcpy(points, outline->points.data(), outline->points.size() * sizeof(PFVector2F));
    if (point_flags && !outline->flags.empty()) memcpy(point_flags, outline->flags.data(), outline->flags.size());
    if (contour_offsets) memcpy(contour_offsets, outline->contour_offsets.data(), outline->contour_offsets.size() * sizeof(uint32_t));
}

/* closed: PFOutlineGetContourCount flags. */
void PFOutlineCopyClosed(PFOutlineRef outline, uint8_t *closed) {
    if (outline && closed && !outline->closed.empty()) memcpy(closed, outline->closed.data(), outline->closed.size());
}

void PFOutlineDestroy(PFOutlineRef outline) { delete outline; }

} // extern "C"
