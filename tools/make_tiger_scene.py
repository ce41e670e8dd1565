#!/usr/bin/env python3
"""Builds tests/golden/tiger.npz — the Ghostscript tiger as a FlatScene (SVG user units).

Run in the build container only (needs /root/reference):
    python tools/make_tiger_scene.py

What it restates (scene *inputs* are "parity unpinned": usvg 0.9.1 is not vendored and the reference
holds no test for its SVG front end, SURVEY.md §8c — the committed fixture is the comparison
origin for both the CUDA path and the oracle):
  * svg/src/lib.rs:124-176   one draw path per fill, one per stroke, fill first; strokes use
                             OutlineStrokeToFill with width max(w, HAIRLINE_STROKE_WIDTH = 0.0333)
                             (svg/src/lib.rs:38), butt caps, miter joins (SVG default limit 4)
  * SVG path data            M/m L/l H/h V/v C/c S/s Q/q T/t Z/z -> absolute segments (usvg's job)
  * content/src/stroke.rs:88-445  stroke-to-fill (offset_forward / offset_backward, recursive
                             offsetting with TOLERANCE 0.01, miter joins)
Arithmetic is float64 here and rounded to float32 at the end.
"""
from __future__ import annotations

import math
import os
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathfinder_b200.flat_scene import FILL_RULE_WINDING, SceneBuilderPy  # noqa: E402

SVG = "/root/reference/resources/svg/Ghostscript_Tiger.svg"
HAIRLINE_STROKE_WIDTH = 0.0333
TOLERANCE = 0.01
EPSILON = 0.001  # geometry/src/util.rs:15

_NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?")


def arc_to_cubics(p0, rx, ry, phi_deg, large, sweep, p1):
    """SVG endpoint arc -> cubic Béziers (W3C SVG 1.1 F.6.5/F.6.6 centre parameterisation, then one
    cubic per <= 90 degree piece; usvg does the same normalisation through kurbo)."""
    if p0 == p1:
        return []
    rx, ry = abs(rx), abs(ry)
    if rx == 0 or ry == 0:
        return [("L", p0, p1)]
    phi = math.radians(phi_deg)
    cphi, sphi = math.cos(phi), math.sin(phi)
    dx2, dy2 = (p0[0] - p1[0]) / 2.0, (p0[1] - p1[1]) / 2.0
    x1p, y1p = cphi * dx2 + sphi * dy2, -sphi * dx2 + cphi * dy2
    lam = (x1p * x1p) / (rx * rx) + (y1p * y1p) / (ry * ry)
    scaled_up = lam > 1
    if scaled_up:
        rx, ry = rx * math.sqrt(lam), ry * math.sqrt(lam)
    num = rx * rx * ry * ry - rx * rx * y1p * y1p - ry * ry * x1p * x1p
    den = rx * rx * y1p * y1p + ry * ry * x1p * x1p
    # scaled-up radii: the centre is the chord's midpoint exactly (no rounding residue, see csrc/svg.cpp)
    coef = 0.0 if scaled_up else math.sqrt(max(num / den, 0.0)) * (-1 if large == sweep else 1)
    cxp, cyp = coef * rx * y1p / ry, -coef * ry * x1p / rx
    cx = cphi * cxp - sphi * cyp + (p0[0] + p1[0]) / 2.0
    cy = sphi * cxp + cphi * cyp + (p0[1] + p1[1]) / 2.0

    def angle(ux, uy, vx, vy):
        a = math.atan2(ux * vy - uy * vx, ux * vx + uy * vy)
        return a

    th1 = angle(1, 0, (x1p - cxp) / rx, (y1p - cyp) / ry)
    dth = angle((x1p - cxp) / rx, (y1p - cyp) / ry, (-x1p - cxp) / rx, (-y1p - cyp) / ry)
    if not sweep and dth > 0:
        dth -= 2 * math.pi
    elif sweep and dth < 0:
        dth += 2 * math.pi
    n = max(1, int(math.ceil(abs(dth) / (math.pi / 2) - 1e-9)))
    delta = dth / n
    k = 4.0 / 3.0 * math.tan(delta / 4.0)
    out = []
    cur = p0

    def pt(th):
        x, y = rx * math.cos(th), ry * math.sin(th)
        return (cphi * x - sphi * y + cx, sphi * x + cphi * y + cy)

    def deriv(th):
        x, y = -rx * math.sin(th), ry * math.cos(th)
        return (cphi * x - sphi * y, sphi * x + cphi * y)

    for i in range(n):
        a0, a1 = th1 + i * delta, th1 + (i + 1) * delta
        e = p1 if i == n - 1 else pt(a1)
        d0, d1 = deriv(a0), deriv(a1)
        c0 = (cur[0] + k * d0[0], cur[1] + k * d0[1])
        c1 = (e[0] - k * d1[0], e[1] - k * d1[1])
        out.append(("C", cur, c0, c1, e))
        cur = e
    return out


def parse_path_data(d: str):
    """Yields contours as (segments, closed); a segment is ('L', p0, p1) | ('Q', p0, c, p1) |
    ('C', p0, c0, c1, p1)."""
    pos = 0
    n_chars = len(d)
    contours = []
    cur = (0.0, 0.0)
    start = (0.0, 0.0)
    segs = None
    last_ctrl = None
    last_cmd = None
    cmd = None

    def skip():
        nonlocal pos
        while pos < n_chars and d[pos] in " \t\r\n,":
            pos += 1

    def num():
        nonlocal pos
        skip()
        m = _NUM.match(d, pos)
        assert m, f"expected number at {pos}: {d[pos:pos + 20]!r}"
        pos = m.end()
        return float(m.group(0))

    def flag():
        nonlocal pos
        skip()
        assert d[pos] in "01", f"expected arc flag at {pos}"
        pos += 1
        return d[pos - 1] == "1"

    def flush(closed):
        nonlocal segs
        if segs is not None:
            contours.append((segs, closed))
        segs = None

    while True:
        skip()
        if pos >= n_chars:
            break
        if d[pos].isalpha():
            cmd = d[pos]
            pos += 1
            if cmd in "Zz":
                flush(True)
                cur = start
                last_cmd = "Z"
                continue
        elif cmd in "Mm":  # implicit lineto after moveto
            cmd = "L" if cmd == "M" else "l"
        rel = cmd.islower()
        c = cmd.upper()
        ox, oy = cur if rel else (0.0, 0.0)
        if c == "M":
            flush(False)
            cur = (ox + num(), oy + num())
            start = cur
            segs = []
        else:
            if segs is None:  # drawing after Z without M: new subpath at the old start
                segs = []
                start = cur
            if c == "L":
                p = (ox + num(), oy + num())
                segs.append(("L", cur, p))
                cur = p
            elif c == "H":
                p = ((cur[0] if rel else 0.0) + num(), cur[1])
                segs.append(("L", cur, p))
                cur = p
            elif c == "V":
                p = (cur[0], (cur[1] if rel else 0.0) + num())
                segs.append(("L", cur, p))
                cur = p
            elif c == "C":
                c0 = (ox + num(), oy + num())
                c1 = (ox + num(), oy + num())
                p = (ox + num(), oy + num())
                segs.append(("C", cur, c0, c1, p))
                last_ctrl = c1
                cur = p
            elif c == "S":
                if last_cmd in ("C", "S") and last_ctrl is not None:
                    c0 = (2 * cur[0] - last_ctrl[0], 2 * cur[1] - last_ctrl[1])
                else:
                    c0 = cur
                c1 = (ox + num(), oy + num())
                p = (ox + num(), oy + num())
                segs.append(("C", cur, c0, c1, p))
                last_ctrl = c1
                cur = p
            elif c == "Q":
                q = (ox + num(), oy + num())
                p = (ox + num(), oy + num())
                segs.append(("Q", cur, q, p))
                last_ctrl = q
                cur = p
            elif c == "T":
                if last_cmd in ("Q", "T") and last_ctrl is not None:
                    q = (2 * cur[0] - last_ctrl[0], 2 * cur[1] - last_ctrl[1])
                else:
                    q = cur
                p = (ox + num(), oy + num())
                segs.append(("Q", cur, q, p))
                last_ctrl = q
                cur = p
            elif c == "A":
                rx, ry, rot = num(), num(), num()
                large, sweep = flag(), flag()
                p = (ox + num(), oy + num())
                segs.extend(arc_to_cubics(cur, rx, ry, rot, large, sweep, p))
                cur = p
            else:
                raise NotImplementedError(f"path command {cmd}")
        last_cmd = c
    flush(False)
    return [(s, closed) for s, closed in contours if s]


# ---- stroke-to-fill (content/src/stroke.rs) -------------------------------------------------------

def _sub(a, b): return (a[0] - b[0], a[1] - b[1])
def _add(a, b): return (a[0] + b[0], a[1] + b[1])
def _mul(a, s): return (a[0] * s, a[1] * s)
def _sqlen(a): return a[0] * a[0] + a[1] * a[1]
def _lerp(a, b, t): return (a[0] + (b[0] - a[0]) * t, a[1] + (b[1] - a[1]) * t)


def _norm(a):
    l = math.sqrt(_sqlen(a))
    return (a[0] / l, a[1] / l)


def line_offset(p0, p1, distance):
    """LineSegment2F::offset (geometry/src/line_segment.rs:241-247)."""
    v = _sub(p1, p0)
    if v == (0.0, 0.0):
        return p0, p1
    n = _norm((v[1], v[0]))
    o = (n[0] * -distance, n[1] * distance)
    return _add(p0, o), _add(p1, o)


def intersection_t(a0, a1, b0, b1):
    """LineSegment2F::intersection_t (line_segment.rs:219-229)."""
    p0p1 = _sub(a1, a0)
    ov = _sub(b1, b0)
    # matrix = [ov.x, ov.y, -p0p1.x, -p0p1.y] column-major -> m11=ov.x m21=ov.y m12=-p0p1.x m22=-p0p1.y
    m11, m21, m12, m22 = ov[0], ov[1], -p0p1[0], -p0p1[1]
    det = m11 * m22 - m12 * m21
    if abs(det) < 0.0001:
        return None
    r = _sub(a0, b0)
    # inverse * r, take y
    inv_det = 1.0 / det
    y = (-m21 * r[0] + m11 * r[1]) * inv_det
    return y


def seg_kind(s): return s[0]
def seg_from(s): return s[1]
def seg_to(s): return s[-1]


def seg_to_cubic(s):
    if s[0] == "C":
        return s
    if s[0] == "Q":
        p0, c, p1 = s[1], s[2], s[3]
        c2 = _add(c, c)
        return ("C", p0, _mul(_add(p0, c2), 1.0 / 3.0), _mul(_add(c2, p1), 1.0 / 3.0), p1)
    raise ValueError


def seg_split(s, t):
    if s[0] == "L":
        m = _lerp(s[1], s[2], t)
        return ("L", s[1], m), ("L", m, s[2])
    _, p0, p1, p2, p3 = seg_to_cubic(s)
    p01, p12, p23 = _lerp(p0, p1, t), _lerp(p1, p2, t), _lerp(p2, p3, t)
    p012, p123 = _lerp(p01, p12, t), _lerp(p12, p23, t)
    p0123 = _lerp(p012, p123, t)
    return ("C", p0, p01, p012, p0123), ("C", p0123, p123, p23, p3)


def seg_sample(s, t):
    if s[0] == "L":
        return _add(s[1], _mul(_sub(s[2], s[1]), t))
    a, _b = seg_split(s, t)  # CubicSegment::sample = split(t).0.baseline.to()
    return a[4]


def seg_reversed(s):
    if s[0] == "L":
        return ("L", s[2], s[1])
    if s[0] == "Q":
        return ("Q", s[3], s[2], s[1])
    return ("C", s[4], s[3], s[2], s[1])


def offset_once(s, d):
    if s[0] == "L":
        a, b = line_offset(s[1], s[2], d)
        return ("L", a, b)

    def ctrl_of(seg0, seg1):
        t = intersection_t(seg0[0], seg0[1], seg1[0], seg1[1])
        if t is not None:
            return _add(seg0[0], _mul(_sub(seg0[1], seg0[0]), t))
        return _lerp(seg0[1], seg1[0], 0.5)

    if s[0] == "Q":
        s0 = line_offset(s[1], s[2], d)
        s1 = line_offset(s[2], s[3], d)
        return ("Q", s0[0], ctrl_of(s0, s1), s1[1])
    _, p0, c0, c1, p3 = s
    if p0 == c0:
        s0 = line_offset(p0, c1, d)
        s1 = line_offset(c1, p3, d)
        return ("C", s0[0], s0[0], ctrl_of(s0, s1), s1[1])
    if c1 == p3:
        s0 = line_offset(p0, c0, d)
        s1 = line_offset(c0, p3, d)
        return ("C", s0[0], ctrl_of(s0, s1), s1[1], s1[1])
    s0 = line_offset(p0, c0, d)
    s1 = line_offset(c0, c1, d)
    s2 = line_offset(c1, p3, d)
    t0 = intersection_t(s0[0], s0[1], s1[0], s1[1])
    t1 = intersection_t(s1[0], s1[1], s2[0], s2[1])
    if t0 is not None and t1 is not None:
        k0 = _add(s0[0], _mul(_sub(s0[1], s0[0]), t0))
        k1 = _add(s1[0], _mul(_sub(s1[1], s1[0]), t1))
    else:
        k0 = _lerp(s0[1], s1[0], 0.5)
        k1 = _lerp(s1[1], s2[0], 0.5)
    return ("C", s0[0], k0, k1, s2[1])


def error_within_tolerance(s, other, d):
    mn, mx = abs(d) - TOLERANCE, abs(d) + TOLERANCE
    mn = 0.0 if mn <= 0 else mn * mn
    mx = 0.0 if mx <= 0 else mx * mx
    for i in range(17):
        t = i / 16.0
        sq = _sqlen(_sub(seg_sample(s, t), seg_sample(other, t)))
        if sq < mn or sq > mx:
            return False
    return True


class OutContour:
    def __init__(self):
        self.points = []
        self.flags = []

    def push_endpoint(self, p):
        self.points.append(p)
        self.flags.append(0)

    def push_segment(self, s):
        self.points.append(s[1]); self.flags.append(0)
        if s[0] == "Q":
            self.points.append(s[2]); self.flags.append(1)
        elif s[0] == "C":
            self.points.append(s[2]); self.flags.append(1)
            self.points.append(s[3]); self.flags.append(2)
        self.points.append(s[-1]); self.flags.append(0)

    def might_need_join(self):
        return len(self.points) >= 2  # Miter

    def add_join(self, distance, miter_limit, join_point, next_tangent):
        p0, p1 = self.points[-2], self.points[-1]
        if _sqlen(_sub(p1, p0)) < EPSILON or _sqlen(_sub(next_tangent[1], next_tangent[0])) < EPSILON:
            return
        t = intersection_t(p0, p1, next_tangent[0], next_tangent[1])
        if t is None or t < -EPSILON:
            return
        miter_endpoint = _add(p0, _mul(_sub(p1, p0), t))
        threshold = miter_limit * distance
        if _sqlen(_sub(miter_endpoint, join_point)) > threshold * threshold:
            return
        self.push_endpoint(miter_endpoint)


def seg_offset(s, d, use_join, miter_limit, out: OutContour):
    """Offset::offset (stroke.rs:252-274)."""
    join_point = seg_from(s)

    def add_to_contour(c):
        if use_join and out.might_need_join():
            p3 = seg_from(c)
            p4 = seg_to(c) if c[0] == "L" else c[2]
            out.add_join(d, miter_limit, join_point, (p4, p3))
        out.push_segment(c)

    if _sqlen(_sub(seg_to(s), seg_from(s))) < TOLERANCE * TOLERANCE:
        add_to_contour(s)
        return
    cand = offset_once(s, d)
    if error_within_tolerance(s, cand, d):
        add_to_contour(cand)
        return
    a, b = seg_split(s, 0.5)
    seg_offset(a, d, use_join, miter_limit, out)
    seg_offset(b, d, use_join, miter_limit, out)


def contour_segments(segs, closed):
    """ContourIter (outline.rs:1014-1074) over the contour Outline::from_segments builds: the
    listed segments plus, for closed contours, the implicit closing line."""
    out = list(segs)
    if closed:
        out.append(("L", seg_to(segs[-1]), seg_from(segs[0])))
    return out


def stroke_to_fill(contours, width, miter_limit=4.0):
    """OutlineStrokeToFill::offset (stroke.rs:88-131). Returns output contours (points, flags)."""
    radius = width * 0.5
    result = []
    for segs, closed in contours:
        all_segs = contour_segments(segs, closed)
        out = OutContour()
        for i, s in enumerate(all_segs):  # offset_forward
            seg_offset(s, -radius, i != 0, miter_limit, out)

        def push(o, was_closed, input_first_point):
            if was_closed and o.might_need_join():
                p1, p0 = o.points[1], o.points[0]
                o.add_join(radius, miter_limit, input_first_point, (p1, p0))
            result.append(o)

        first_point = seg_from(segs[0])
        if closed:
            push(out, True, first_point)
            out = OutContour()
        # butt caps: add_cap is a no-op
        rev = [seg_reversed(s) for s in all_segs][::-1]
        for i, s in enumerate(rev):  # offset_backward
            seg_offset(s, -radius, i != 0, miter_limit, out)
        push(out, closed, first_point)
    return result


def parse_color(s: str):
    s = s.strip()
    assert s.startswith("#"), s
    h = s[1:]
    if len(h) == 3:
        h = "".join(ch * 2 for ch in h)
    return (int(h[0:2], 16), int(h[2:4], 16), int(h[4:6], 16), 255)


def main():
    tree = ET.parse(SVG)
    root = tree.getroot()
    ns = "{http://www.w3.org/2000/svg}"
    vb = [float(v) for v in root.attrib["viewBox"].split()]
    builder = SceneBuilderPy((vb[0], vb[1], vb[0] + vb[2], vb[1] + vb[3]))
    n_fill = n_stroke = 0
    for g in root.iter(ns + "g"):
        inherited_fill = g.attrib.get("fill")
        for el in g.findall(ns + "path"):
            contours = parse_path_data(el.attrib["d"])
            fill = el.attrib.get("fill", inherited_fill)
            stroke = el.attrib.get("stroke")
            if fill and fill != "none":
                for segs, _closed in contours:
                    builder.move_to(*seg_from(segs[0]))
                    for s in segs:
                        if s[0] == "L":
                            builder.line_to(*s[2])
                        elif s[0] == "Q":
                            builder.quad_to(*s[2], *s[3])
                        else:
                            builder.cubic_to(*s[2], *s[3], *s[4])
                    builder.close()
                builder.end_path(parse_color(fill), FILL_RULE_WINDING)
                n_fill += 1
            if stroke and stroke != "none":
                width = max(float(el.attrib.get("stroke-width", "1")), HAIRLINE_STROKE_WIDTH)
                for o in stroke_to_fill(contours, width):
                    if not o.points:
                        continue
                    builder._end_contour()
                    builder._points += o.points
                    builder._flags += o.flags
                    builder._open = True
                    builder._end_contour()
                builder.end_path(parse_color(stroke), FILL_RULE_WINDING)
                n_stroke += 1
    scene = builder.finish("tiger")
    out = os.path.join(ROOT, "tests", "golden", "tiger.npz")
    if "--check" in sys.argv[1:]:  # compare with the committed fixture, write nothing
        from pathfinder_b200.flat_scene import FlatScene
        have = FlatScene.load(out)
        same = all(np.array_equal(getattr(have, k), getattr(scene, k)) for k in
                   ("points", "point_flags", "contour_offsets", "path_contour_offsets", "fill_rules", "paints", "paint_colors"))
        print("tiger.npz:", "identical" if same else "DIFFERENT")
        sys.exit(0 if same else 1)
    scene.save(out)
    print(f"{n_fill} fills + {n_stroke} strokes = {scene.n_paths} paths, {scene.n_contours} contours, "
          f"{len(scene.points)} points -> {out} ({os.path.getsize(out)} bytes)")


if __name__ == "__main__":
    main()
