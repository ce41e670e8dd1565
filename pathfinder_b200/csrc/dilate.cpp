// pathfinder_b200/csrc/dilate.cpp — Outline::dilate on the host (stem darkening).
//
// Restates content/src/outline.rs:243-249 (Outline::dilate), content/src/orientation.rs:28-57
// (Orientation::from_outline) and content/src/dilation.rs:34-125 (ContourDilator::dilate) in scalar f32 with the
// reference's operation order, so that a scene prepared here and the CPU tiler's apply_render_options
// (renderer/src/scene.rs:268-270) hand bit-identical points to the tiler. Every distinct position of a contour
// moves along the bisector of its two neighbouring edge directions, by `amount` per axis; runs of coincident
// points move together.
#include "dilate.h"

#include <cmath>

namespace pf {

namespace {

struct Vec {
    float x, y;
};
inline Vec sub(PFVector2F a, PFVector2F b) { return Vec{a.x - b.x, a.y - b.y}; }
inline bool same(PFVector2F a, PFVector2F b) { return a.x == b.x && a.y == b.y; }
// Vector2F::normalize: self * (1.0 / self.length()) (geometry/src/vector.rs)
inline Vec unit(Vec v) {
    const float k = 1.0f / sqrtf(v.x * v.x + v.y * v.y);
    return Vec{v.x * k, v.y * k};
}

} // namespace

bool outline_is_clockwise(const PFVector2F *points, const uint32_t *contour_offsets, uint32_t contour_count) {
    // Twice the signed area, accumulated in f32 in point order: sum of det(previous, next) around every contour,
    // starting from the contour's last point. Positive = clockwise with y down; zero counts as counterclockwise.
    float area = 0.0f;
    for (uint32_t c = 0; c < contour_count; c++) {
        const uint32_t begin = contour_offsets[c], end = contour_offsets[c + 1];
        if (begin == end) continue;
        PFVector2F before = points[end - 1];
        for (uint32_t i = begin; i < end; i++) {
            area += before.x * points[i].y - before.y * points[i].x;
            before = points[i];
        }
    }
    return area > 0.0f;
}

void dilate_contour(PFVector2F *pts, uint32_t n, float amount_x, float amount_y, bool clockwise) {
    if (n == 0) return;
    // amount * (1, -1) for counterclockwise outlines, amount * (-1, 1) for clockwise ones (dilation.rs:36-39).
    const float kx = clockwise ? amount_x * -1.0f : amount_x * 1.0f;
    const float ky = clockwise ? amount_y * 1.0f : amount_y * -1.0f;

    // The walk starts at point 0's position; `tail` is the last point (going backwards from the end) that is not
    // at that position, so [tail + 1, n) + [0, ...) is the first run of coincident points.
    const PFVector2F origin = pts[0];
    uint32_t tail = 0;
    do {
        tail = tail == 0 ? n - 1 : tail - 1;
    } while (tail != 0 && same(pts[tail], origin));
    const uint32_t run0 = tail + 1 == n ? 0 : tail + 1; // first index of the first run
    Vec incoming = unit(sub(origin, pts[tail]));

    uint32_t run = run0;
    PFVector2F here = origin;
    for (;;) {
        // The run [run, after) of points at `here`, and the next distinct position. Wrapping round to the first
        // run ends the walk; its position is the ORIGINAL point 0, which by then has been overwritten in pts.
        uint32_t after = run;
        PFVector2F there;
        bool wrapped = false;
        for (;;) {
            after = after + 1 == n ? 0 : after + 1;
            if (after == run0) {
                there = origin;
                wrapped = true;
                break;
            }
            there = pts[after];
            if (after == run || !same(there, here)) break;
        }
        const Vec outgoing = unit(sub(there, here));

        // bisector = incoming.yx() + outgoing.yx(); the point moves by -(bisector * scale) / |bisector|.
        const float bx = incoming.y + outgoing.y, by = incoming.x + outgoing.x;
        const float len = sqrtf(bx * bx + by * by);
        float mx = 0.0f, my = 0.0f;
        if (len != 0.0f) {
            const float inv = 1.0f / len;
            mx = (bx * kx) * inv;
            my = (by * ky) * inv;
        }
        const PFVector2F moved{here.x - mx, here.y - my};
        for (uint32_t i = run; i != after; i = i + 1 == n ? 0 : i + 1) pts[i] = moved;

        if (wrapped) break;
        incoming = outgoing;
        here = there;
        run = after;
    }
}

void dilate_outline(PFVector2F *points, const uint32_t *contour_offsets, uint32_t contour_count, float amount_x,
                    float amount_y) {
    const bool clockwise = outline_is_clockwise(points, contour_offsets, contour_count);
    for (uint32_t c = 0; c < contour_count; c++)
        dilate_contour(points + contour_offsets[c], contour_offsets[c + 1] - contour_offsets[c], amount_x, amount_y,
                       clockwise);
}

} // namespace pf

extern "C" void PFOutlineDilate(PFVector2F *points, const uint32_t *contour_offsets, uint32_t contour_count,
                                const PFVector2F *amount) {
    if (!points || !contour_offsets || !amount || contour_count == 0) return;
    if (amount->x == 0.0f && amount->y == 0.0f) return; // Vector2F::is_zero: callers skip the pass (scene.rs:268)
    pf::dilate_outline(points, contour_offsets, contour_count, amount->x, amount->y);
}
