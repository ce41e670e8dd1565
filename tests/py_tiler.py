"""A second, independent restatement of the reference's CPU tiler in scalar numpy float32 — test infrastructure.

oracle/pf_oracle.cpp is "parity unpinned": nothing in this image can run the Rust tiler. This module narrows the room
for a transcription error by restating the same functions a second time, straight from the Rust sources and without
looking at the C++, in the slowest and plainest form available (one np.float32 operation per Rust operation, Python
loops), so that tests/test_py_tiler.py can demand bit-identical fills, alpha-tile ids and backdrops from both on
small scenes. Restated here:
  ContourIter::next                      content/src/outline.rs:1019-1062
  Segment::to_cubic                      content/src/segment.rs:171-183
  CubicSegment::is_flat / split          content/src/segment.rs:292-360
  process_segment / process_line_segment renderer/src/tiler.rs:166-308
  clip_line_segment_to_rect              content/src/clip.rs:494-565
  ObjectBuilder::add_fill, adjust_alpha_tile_backdrop, get_or_allocate_alpha_tile_index
                                         renderer/src/builder.rs:509-616
  Tiler::new bounds, round_rect_out_to_tile_bounds, prepare_tiles (with the four clip cases)
                                         renderer/src/tiler.rs:47-50,100-165; renderer/src/tiles.rs:64-66
plus the D3D9 batch pack with its z-buffer (renderer/src/builder.rs:1011-1040). Solid-colour draw paths, clip paths one
level deep, no transform; sequential order (SequentialExecutor)."""
from __future__ import annotations

import numpy as np

f = np.float32
TILE = 16
INVALID = 0xFFFFFFFF


def contour_segments(pts, flags):
    """ContourIter with the close segment of a closed contour: ('line' | 'quad' | 'cubic', points...)."""
    n = len(pts)
    out, index = [], 1
    while True:
        if index == n + 1:
            break
        p0 = pts[index - 1]
        if index == n:
            out.append(("line", p0, pts[0]))
            index += 1
            continue
        p1 = pts[index]
        index += 1
        if flags[index - 1] == 0:
            out.append(("line", p0, p1))
            continue
        p2 = pts[index]
        index += 1
        if flags[index - 1] == 0:
            out.append(("quad", p0, p1, p2))
            continue
        p3 = pts[index]
        index += 1
        out.append(("cubic", p0, p1, p2, p3))
    return out


def add(a, b):
    return (a[0] + b[0], a[1] + b[1])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1])


def scale(a, k):
    return (a[0] * k, a[1] * k)


def lerp_pt(a, b, t):  # a + t * (b - a), per component
    return (a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]))


def is_flat(p0, c0, c1, p3):
    three = f(3.0)
    uv = [three * c0[0] - p0[0] - p0[0] - p3[0], three * c0[1] - p0[1] - p0[1] - p3[1],
          three * c1[0] - p3[0] - p3[0] - p0[0], three * c1[1] - p3[1] - p3[1] - p0[1]]
    uv = [v * v for v in uv]
    m0 = uv[0] if uv[0] > uv[2] else uv[2]   # uv.max(uv.zwxy())
    m1 = uv[1] if uv[1] > uv[3] else uv[3]
    return m0 + m1 <= f(16.0) * f(0.25) * f(0.25)


def split_half(p0, c0, c1, p3):
    t = f(0.5)
    p01, p12, p23 = lerp_pt(p0, c0, t), lerp_pt(c0, c1, t), lerp_pt(c1, p3, t)
    p012, p123 = lerp_pt(p01, p12, t), lerp_pt(p12, p23, t)
    p0123 = lerp_pt(p012, p123, t)
    return (p0, p01, p012, p0123), (p0123, p123, p23, p3)


class PathBuilder:
    """ObjectBuilder for one path: dense tile map over the tile rect, per-column backdrops above it."""

    def __init__(self, bounds, view_box, state):
        self.view_box = view_box
        self.state = state  # shared: {"next_alpha": int, "fills": list}
        b = intersection(bounds, view_box)
        k = f(1.0) / f(TILE)
        self.x0, self.y0 = int(np.floor(b[0] * k)), int(np.floor(b[1] * k))
        self.x1, self.y1 = int(np.ceil(b[2] * k)), int(np.ceil(b[3] * k))
        self.w, self.h = self.x1 - self.x0, self.y1 - self.y0
        self.alpha = {}      # (tx, ty) -> alpha tile id
        self.backdrop = {}   # (tx, ty) -> i8 delta, then the propagated backdrop
        self.col_backdrop = [0] * max(self.w, 0)

    def inside(self, tx, ty):
        return self.x0 <= tx < self.x1 and self.y0 <= ty < self.y1

    def add_fill(self, frm, to, tx, ty):
        if not self.inside(tx, ty):
            return
        ox, oy = f(tx) * f(TILE), f(ty) * f(TILE)
        vals = [(frm[0] - ox) * f(256.0), (frm[1] - oy) * f(256.0), (to[0] - ox) * f(256.0), (to[1] - oy) * f(256.0)]
        hi = f(TILE * 256 - 1)
        q = []
        for v in vals:
            v = v if v > f(0.0) else f(0.0)       # clamp: max(min), then min(max)
            v = v if v < hi else hi
            q.append(int(np.rint(v)))             # cvtps: round to nearest, ties to even
        if q[0] == q[2]:
            return
        key = (tx, ty)
        if key not in self.alpha:
            self.alpha[key] = self.state["next_alpha"]
            self.state["next_alpha"] += 1
        self.state["fills"].append((q[0], q[1], q[2], q[3], self.alpha[key]))

    def adjust_backdrop(self, tx, ty, delta):
        ox, oy = tx - self.x0, ty - self.y0
        if ox < 0 or ox >= self.w or oy >= self.h:
            return
        if oy < 0:
            self.col_backdrop[ox] += delta
            return
        self.backdrop[(tx, ty)] = self.backdrop.get((tx, ty), 0) + delta

    def prepare_tiles(self, clip=None):
        """Tiler::prepare_tiles (tiler.rs:100-165). `clip` = (tile rect, prepared tiles) of the built clip path, or
        None. Returns ({(tx, ty): (alpha tile id | INVALID, backdrop)} for every tile of the rect, clip records in
        row-major order as (dest_tile_id, dest_backdrop, src_tile_id, src_backdrop))."""
        out, clips = {}, []
        cols = list(self.col_backdrop)
        for ty in range(self.y0, self.y1):
            for tx in range(self.x0, self.x1):
                col = tx - self.x0
                delta = self.backdrop.get((tx, ty), 0)
                alpha, backdrop = self.alpha.get((tx, ty), INVALID), ((cols[col] + 128) % 256) - 128
                if clip is not None:
                    (cx0, cy0, cx1, cy1), clip_tiles = clip
                    if cx0 <= tx < cx1 and cy0 <= ty < cy1:
                        clip_alpha, clip_backdrop = clip_tiles[(tx, ty)]
                        if clip_alpha != INVALID and alpha != INVALID:
                            clips.append((alpha, backdrop, clip_alpha, clip_backdrop))  # combine the two masks
                            backdrop = 0
                        elif clip_alpha != INVALID and alpha == INVALID and backdrop != 0:
                            alpha, backdrop = clip_alpha, clip_backdrop                 # solid tile: the clip's mask
                        elif clip_alpha == INVALID and clip_backdrop == 0:
                            alpha, backdrop = INVALID, 0                                # blank clip tile: cull
                    else:
                        alpha, backdrop = INVALID, 0                                    # outside the clip's rect
                out[(tx, ty)] = (alpha, backdrop)
                cols[col] += delta
        return out, clips


def intersection(a, b):
    """RectF::intersection(...).unwrap_or(RectF::default()) (geometry/src/rect.rs:122-137)."""
    if not (a[0] < b[2] and a[1] < b[3] and b[0] < a[2] and b[1] < a[3]):
        return (f(0), f(0), f(0), f(0))
    return (max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3]))


def outcode(p, r):
    code = 0
    if p[0] < r[0]:
        code |= 1   # LEFT
    if p[1] < r[1]:
        code |= 4   # TOP
    if p[0] > r[2]:
        code |= 2   # RIGHT
    if p[1] > r[3]:
        code |= 8   # BOTTOM
    return code


def lerp(a, b, t):  # util.rs: a + (b - a) * t
    return a + (b - a) * t


def clip_line(frm, to, r):
    cf, ct = outcode(frm, r), outcode(to, r)
    while True:
        if cf == 0 and ct == 0:
            return frm, to
        if cf & ct:
            return None
        clip_from = cf > ct
        code = cf if clip_from else ct
        if code & 1:
            p = (r[0], lerp(frm[1], to[1], (r[0] - frm[0]) / (to[0] - frm[0])))
        elif code & 2:
            p = (r[2], lerp(frm[1], to[1], (r[2] - frm[0]) / (to[0] - frm[0])))
        elif code & 4:
            p = (lerp(frm[0], to[0], (r[1] - frm[1]) / (to[1] - frm[1])), r[1])
        else:
            p = (lerp(frm[0], to[0], (r[3] - frm[1]) / (to[1] - frm[1])), r[3])
        if clip_from:
            frm, cf = p, outcode(p, r)
        else:
            to, ct = p, outcode(p, r)


def process_line_segment(frm, to, b: PathBuilder, lines=None):
    if lines is not None:
        lines.append((frm[0], frm[1], to[0], to[1]))
    vb = b.view_box
    with np.errstate(all="ignore"):
        clipped = clip_line(frm, to, (vb[0], f(-np.inf), vb[2], vb[3]))
        if clipped is None:
            return
        frm, to = clipped
        k = f(1.0) / f(TILE)
        ftx, fty = int(np.floor(frm[0] * k)), int(np.floor(frm[1] * k))
        ttx, tty = int(np.floor(to[0] * k)), int(np.floor(to[1] * k))
        vx, vy = to[0] - frm[0], to[1] - frm[1]
        neg_x, neg_y = bool(vx < 0), bool(vy < 0)
        step_x, step_y = (-1 if neg_x else 1), (-1 if neg_y else 1)
        cross_x = f(ftx + (0 if neg_x else 1)) * f(TILE)
        cross_y = f(fty + (0 if neg_y else 1)) * f(TILE)
        t_max_x, t_max_y = (cross_x - frm[0]) / vx, (cross_y - frm[1]) / vy
        t_delta_x, t_delta_y = abs(f(TILE) / vx), abs(f(TILE) / vy)
        cur = frm
        tx, ty = ftx, fty
        last = None
        while True:
            if t_max_x < t_max_y:
                nxt = "x"
            elif t_max_x > t_max_y:
                nxt = "y"
            else:
                nxt = "x" if step_x > 0 else "y"
            next_t = t_max_x if nxt == "x" else t_max_y
            next_t = next_t if next_t < f(1.0) else f(1.0)   # f32::min(next_t, 1.0); NaN -> 1.0
            if not (next_t == next_t):
                next_t = f(1.0)
            if (tx, ty) == (ttx, tty):
                nxt = None
            nxt_pos = (frm[0] + vx * next_t, frm[1] + vy * next_t)
            b.add_fill(cur, nxt_pos, tx, ty)
            corner = (f(tx) * f(TILE), f(ty) * f(TILE))
            if step_y < 0 and nxt == "y":
                b.add_fill(nxt_pos, corner, tx, ty)
            elif step_y > 0 and last == "y":
                b.add_fill(corner, cur, tx, ty)
            if step_x < 0 and last == "x":
                b.adjust_backdrop(tx, ty, 1)
            elif step_x > 0 and nxt == "x":
                b.adjust_backdrop(tx, ty, -1)
            if nxt is None:
                break
            if nxt == "x":
                if tx == ttx:
                    break
                t_max_x = t_max_x + t_delta_x
                tx += step_x
            else:
                if ty == tty:
                    break
                t_max_y = t_max_y + t_delta_y
                ty += step_y
            cur = nxt_pos
            last = nxt


def process_cubic(p0, c0, c1, p3, b, lines):
    if is_flat(p0, c0, c1, p3):
        process_line_segment(p0, p3, b, lines)
        return
    first, second = split_half(p0, c0, c1, p3)
    process_cubic(*first, b, lines)
    process_cubic(*second, b, lines)


def process_segment(seg, b, lines):
    if seg[0] == "line":
        process_line_segment(seg[1], seg[2], b, lines)
    elif seg[0] == "quad":
        p0, c, p3 = seg[1], seg[2], seg[3]
        c2 = add(c, c)
        third = f(1.0) / f(3.0)
        process_cubic(p0, scale(add(p0, c2), third), scale(add(c2, p3), third), p3, b, lines)
    else:
        process_cubic(seg[1], seg[2], seg[3], seg[4], b, lines)


def _tile_outline(flat, c0, c1, vb, state, clip=None):
    co = flat.contour_offsets
    contours = []
    for c in range(int(c0), int(c1)):
        pts = [(f(x), f(y)) for x, y in flat.points[int(co[c]):int(co[c + 1])]]
        if pts:
            contours.append((pts, [int(v) for v in flat.point_flags[int(co[c]):int(co[c + 1])]]))
    allp = [q for pts, _ in contours for q in pts]
    if allp:
        bounds = (min(q[0] for q in allp), min(q[1] for q in allp), max(q[0] for q in allp), max(q[1] for q in allp))
    else:
        bounds = (f(0), f(0), f(0), f(0))
    b = PathBuilder(bounds, vb, state)
    path_lines = []
    for pts, fl in contours:
        for seg in contour_segments(pts, fl):
            process_segment(seg, b, path_lines)
    tiles, clips = b.prepare_tiles(clip)
    return (b.x0, b.y0, b.x1, b.y1), tiles, clips, path_lines


def tile_scene(flat):
    """Tiles a FlatScene (no transform; clip paths one level deep) the way SceneBuilder::build does on the CPU with a
    SequentialExecutor: every clip path first, then the draw paths (builder.rs:224-325). Returns a dict:
      fills       [(from_x, from_y, to_x, to_y, alpha tile id)] in emission order
      lines       flattened segments per path, clip paths first
      tiles       the D3D9 batch's tile list (builder.rs:1011-1028): (tile_x, tile_y, alpha id, draw path id, backdrop)
                  of every tile with a mask or a backdrop, path by path, row-major
      clips       the batch's Clip records (builder.rs:1031-1040)
      z_buffer    max draw path id over solid tiles of occluding paths, over the view box's tile rect, initially 0"""
    state = {"next_alpha": 0, "fills": []}
    vb = tuple(f(v) for v in flat.view_box)
    lines, built_clips = [], []
    for (c0, c1) in flat.clip_contour_ranges:
        rect, tiles, _, path_lines = _tile_outline(flat, c0, c1, vb, state)
        built_clips.append((rect, tiles))
        lines.append(path_lines)
    k = f(1.0) / f(TILE)
    zx0, zy0 = int(np.floor(vb[0] * k)), int(np.floor(vb[1] * k))
    zx1, zy1 = int(np.ceil(vb[2] * k)), int(np.ceil(vb[3] * k))
    z = np.zeros((max(zy1 - zy0, 0), max(zx1 - zx0, 0)), np.int32)
    batch_tiles, batch_clips = [], []
    paints, colors = flat.palette()
    for p in range(flat.n_paths):
        clip_id = int(flat.draw_clip_paths[p])
        clip = None if clip_id == INVALID else built_clips[clip_id]
        rect, tiles, clips, path_lines = _tile_outline(flat, flat.path_contour_offsets[p], flat.path_contour_offsets[p + 1],
                                                       vb, state, clip)
        lines.append(path_lines)
        occludes = int(colors[int(paints[p])][3]) == 255   # BuiltDrawPath::new: opaque paint and SrcOver
        for ty in range(rect[1], rect[3]):
            for tx in range(rect[0], rect[2]):
                alpha, backdrop = tiles[(tx, ty)]
                if alpha == INVALID and backdrop == 0:
                    continue
                batch_tiles.append((tx, ty, alpha, p, backdrop))
                if occludes and alpha == INVALID:
                    z[ty - zy0, tx - zx0] = max(z[ty - zy0, tx - zx0], p)
        batch_clips += [c for c in clips if c[0] != INVALID and c[2] != INVALID]
    return {"fills": state["fills"], "lines": lines, "tiles": batch_tiles, "clips": batch_clips, "z_buffer": z,
            "z_rect": (zx0, zy0, zx1, zy1)}


# ---------------------------------------------------------------------------------------------------------------
# Coverage and compositing, restated from the shaders (floating point: compared within a tolerance, not bit for bit).
#   computeCoverage                 shaders/fill_area.inc.glsl:11-27
#   accumulateCoverageForFillList   shaders/d3d11/fill_compute.inc.glsl:11-25
#   clip combine                    shaders/d3d9/tile_clip_combine.fs.glsl:28-31
#   sampleMask, calculateColor      shaders/tile_fragment.inc.glsl:539-614
#   dest = dest * (1 - a) + src     shaders/d3d11/tile.cs.glsl:155
# with the decisions of DESIGN.md §2 (D3D9 semantics: unclamped coverage, fill rule at composite, solid tiles use
# coverage = backdrop through the same rule; tiles behind an occluder are skipped: path id < z).
# ---------------------------------------------------------------------------------------------------------------

def sample_lut(lut, u, v):
    """texture(areaLUT, vec2(u, v)) over the (256, 256, 4) RGBA8 table: bilinear, clamp to edge."""
    h, w, _ = lut.shape
    table = lut.astype(np.float32) / f(255.0)
    x, y = u * f(w) - f(0.5), v * f(h) - f(0.5)
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0)[..., None], (y - y0)[..., None]
    xa, xb = np.clip(x0.astype(np.int64), 0, w - 1), np.clip(x0.astype(np.int64) + 1, 0, w - 1)
    ya, yb = np.clip(y0.astype(np.int64), 0, h - 1), np.clip(y0.astype(np.int64) + 1, 0, h - 1)
    top = table[ya, xa] * (f(1.0) - fx) + table[ya, xb] * fx
    bottom = table[yb, xa] * (f(1.0) - fx) + table[yb, xb] * fx
    return top * (f(1.0) - fy) + bottom * fy


def alpha_masks(fills, n_alpha, lut):
    """(n_alpha, 16, 16) float32 coverage sums: thread (x, strip) handles pixel column x, rows 4 strip .. 4 strip + 3."""
    masks = np.zeros((n_alpha, 16, 16), np.float32)
    cx = (np.arange(16, dtype=np.float32) + f(0.5))[None, :]                # (1, 16): tileFragCoord.x
    cy = (np.arange(4, dtype=np.float32) * f(4.0) + f(0.5))[:, None]        # (4, 1):  tileFragCoord.y of each strip
    with np.errstate(all="ignore"):
        for fx, fy, tx, ty, link in fills:
            from_x, from_y = f(fx) / f(256.0) - cx, f(fy) / f(256.0) - cy
            to_x, to_y = f(tx) / f(256.0) - cx, f(ty) / f(256.0) - cy
            from_x, to_x = np.broadcast_to(from_x, (4, 16)), np.broadcast_to(to_x, (4, 16))
            from_y, to_y = np.broadcast_to(from_y, (4, 16)), np.broadcast_to(to_y, (4, 16))
            from_left = from_x < to_x
            lx, ly = np.where(from_left, from_x, to_x), np.where(from_left, from_y, to_y)
            rx, ry = np.where(from_left, to_x, from_x), np.where(from_left, to_y, from_y)
            wx, wy = np.clip(from_x, f(-0.5), f(0.5)), np.clip(to_x, f(-0.5), f(0.5))
            offset = (wx * f(0.5) + wy * f(0.5)) - lx                        # mix(window.x, window.y, 0.5) - left.x
            t = offset / (rx - lx)
            y = ly * (f(1.0) - t) + ry * t                                   # mix(left.y, right.y, t)
            d = (ry - ly) / (rx - lx)
            dx = wx - wy
            tex = sample_lut(lut, (y + f(8.0)) / f(16.0), np.abs(d * dx) / f(16.0))   # (4, 16, 4): channel k = row + k
            cov = tex * dx[..., None]
            cov = np.where((dx == 0)[..., None], f(0.0), cov)                # 0 * NaN: columns outside the segment
            masks[link] += cov.transpose(0, 2, 1).reshape(16, 16)
    return masks


def f16_round(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def render(flat, result, lut, width, height, background=(0.0, 0.0, 0.0, 0.0)):
    """The frame as float32 RGBA (premultiplied), before the RGBA8 store."""
    n_alpha = 1 + max([t[4] for t in result["fills"]], default=-1)
    masks = alpha_masks(result["fills"], n_alpha, lut)
    for dest_id, dest_backdrop, src_id, src_backdrop in result["clips"]:
        masks[dest_id] = np.minimum(np.abs(masks[dest_id] + f(dest_backdrop)), np.abs(masks[src_id] + f(src_backdrop)))
    dest = np.empty((height, width, 4), np.float32)
    dest[:] = np.asarray(background, np.float32)
    paints, colors = flat.palette()
    zx0, zy0, _, _ = result["z_rect"]
    for tx, ty, alpha, path, backdrop in result["tiles"]:
        if path < result["z_buffer"][ty - zy0, tx - zx0]:
            continue
        x0, y0 = tx * TILE, ty * TILE
        x1, y1 = min(x0 + TILE, width), min(y0 + TILE, height)
        if x0 < 0 or y0 < 0 or x1 <= x0 or y1 <= y0:
            continue
        coverage = (masks[alpha] if alpha != INVALID else np.zeros((16, 16), np.float32)) + f(backdrop)
        if int(flat.fill_rules[path]) == 0:
            coverage = np.abs(coverage)
        else:
            m = coverage - f(2.0) * np.floor(coverage / f(2.0))             # mod(coverage, 2.0)
            coverage = f(1.0) - np.abs(f(1.0) - m)
        mask_alpha = np.minimum(f(1.0), coverage)[: y1 - y0, : x1 - x0]
        base = f16_round(colors[int(paints[path])].astype(np.float32) * (f(1.0) / f(255.0)))  # RGBA16F paint texel
        a = base[3] * mask_alpha
        src = np.stack([base[0] * a, base[1] * a, base[2] * a, a], axis=-1)
        dest[y0:y1, x0:x1] = dest[y0:y1, x0:x1] * (f(1.0) - a)[..., None] + src
    return dest
