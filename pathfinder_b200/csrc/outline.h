// pathfinder_b200/csrc/outline.h — the flat outline container behind PFOutlineRef (include/pf_cuda.h), shared by
// the host-side front-end pieces (stroke.cpp, svg.cpp).
#pragma once

#include <cstdint>
#include <vector>

#include "../../include/pf_cuda.h"

struct PFOutline {
    std::vector<PFVector2F> points;
    std::vector<uint8_t> flags;             // content/src/outline.rs PointFlags
    std::vector<uint32_t> contour_offsets{0};
    std::vector<uint8_t> closed;            // per contour
};
