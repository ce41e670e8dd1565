// pathfinder_b200/csrc/dilate.h — Outline::dilate restated for the host side (dilate.cpp).
#pragma once

#include <cstdint>

#include "../../include/pf_cuda.h"

namespace pf {

// Orientation::from_outline (content/src/orientation.rs:28-57): true = clockwise with y down.
bool outline_is_clockwise(const PFVector2F *points, const uint32_t *contour_offsets, uint32_t contour_count);
// ContourDilator::dilate (content/src/dilation.rs:34-125) on one contour, in place.
void dilate_contour(PFVector2F *points, uint32_t point_count, float amount_x, float amount_y, bool clockwise);
// Outline::dilate (content/src/outline.rs:243-249) without the bounds update: contour c covers
// points[contour_offsets[c] .. contour_offsets[c + 1]).
void dilate_outline(PFVector2F *points, const uint32_t *contour_offsets, uint32_t contour_count, float amount_x,
                    float amount_y);

} // namespace pf
