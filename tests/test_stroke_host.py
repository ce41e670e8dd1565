"""Stroke-to-fill on the host: the C++ restatement of content/src/stroke.rs behind
PFOutlineStrokeToFill, checked through geometric properties and against the independent Python restatement that
built the tiger fixture (tools/make_tiger_scene.py, double precision) on the tiger's own 52 strokes."""
import os

import numpy as np
import pytest

from pathfinder_b200 import _lib as L
from pathfinder_b200 import api

TIGER_SVG = "/root/reference/resources/svg/Ghostscript_Tiger.svg"


def polygon_area(points):
    x, y = points[:, 0].astype(np.float64), points[:, 1].astype(np.float64)
    return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def test_open_line_is_a_rectangle():
    pts, flags, offs = api.stroke_to_fill([(10, 10), (50, 10)], [0, 0], [0, 2], [0], line_width=4.0)
    assert list(offs) == [0, 4] and not flags.any()
    assert np.allclose(sorted(map(tuple, pts)), sorted([(10, 8), (50, 8), (50, 12), (10, 12)]))
    assert abs(polygon_area(pts) - 40 * 4) < 1e-3


def test_square_caps_extend_the_line():
    pts, _flags, offs = api.stroke_to_fill([(10, 10), (50, 10)], [0, 0], [0, 2], [0], line_width=4.0, line_cap="square")
    assert list(offs) == [0, 10]
    assert abs(pts[:, 0].min() - 8) < 1e-5 and abs(pts[:, 0].max() - 52) < 1e-5
    assert abs(pts[:, 1].min() - 8) < 1e-5 and abs(pts[:, 1].max() - 12) < 1e-5


@pytest.mark.parametrize("join,outer", [("miter", 14.0 * 14.0), ("bevel", 14.0 * 14.0 - 4 * 0.5 * 2 * 2)])
def test_closed_square_gives_two_contours(join, outer):
    """A closed 10x10 square stroked with width 4: the outer contour is the square grown by 2 (corners mitred, or cut
    off by 2x2 triangles with bevel joins), the inner one the square shrunk by 2."""
    square = [(20, 20), (30, 20), (30, 30), (20, 30)]
    pts, flags, offs = api.stroke_to_fill(square, [0] * 4, [0, 4], [1], line_width=4.0, line_join=join, miter_limit=10.0)
    assert len(offs) == 3 and not flags.any()
    areas = sorted(polygon_area(pts[offs[i]:offs[i + 1]]) for i in range(2))
    assert abs(areas[1] - outer) < 1e-3
    if join == "miter":  # (with bevel joins the inner offset lines cross at the corners: not a simple polygon)
        assert abs(areas[0] - 6.0 * 6.0) < 1e-3


def test_miter_limit_falls_back_to_bevel():
    """A sharp corner whose miter would be longer than miter_limit * radius gets no miter point."""
    corner = [(0, 0), (40, 0), (0, 4)]  # ~5.7 degrees: miter length ~ 20 x radius
    sharp, _, _ = api.stroke_to_fill(corner, [0] * 3, [0, 3], [0], line_width=2.0, line_join="miter", miter_limit=100.0)
    blunt, _, _ = api.stroke_to_fill(corner, [0] * 3, [0, 3], [0], line_width=2.0, line_join="miter", miter_limit=4.0)
    bevel, _, _ = api.stroke_to_fill(corner, [0] * 3, [0, 3], [0], line_width=2.0, line_join="bevel")
    assert len(sharp) == len(blunt) + 2 and np.array_equal(blunt, bevel)  # one miter point per side of the stroke
    assert sharp[:, 0].max() > 55 and blunt[:, 0].max() < 42


def test_curves_stay_within_tolerance_of_the_offset():
    """A quarter circle (one cubic) stroked with width 6: every output on-curve point lies at distance 3 from the
    circle, within the stroker's tolerance (0.01) plus the cubic's own approximation error."""
    k = 0.5522847498 * 50
    pts, flags, offs = api.stroke_to_fill([(50, 0), (50, k), (k, 50), (0, 50)], [0, 1, 2, 0], [0, 4], [0], line_width=6.0)
    assert len(offs) == 2 and flags.any()
    on_curve = pts[flags == 0]
    r = np.hypot(on_curve[:, 0], on_curve[:, 1])
    assert np.all(np.minimum(np.abs(r - 47), np.abs(r - 53)) < 0.05)
    assert abs(polygon_area_flattened(pts, flags) - 0.25 * np.pi * (53 ** 2 - 47 ** 2)) < 1.0


def polygon_area_flattened(pts, flags, steps=32):
    out, i, n = [], 0, len(pts)
    while i < n:
        if flags[i] == 0:
            out.append(pts[i]); i += 1
            continue
        p0 = out[-1]
        ctrl = [pts[i]]
        if flags[i + 1] != 0:
            ctrl.append(pts[i + 1])
        p_end = pts[i + len(ctrl)]
        for s in range(1, steps + 1):
            t = s / steps
            if len(ctrl) == 1:
                out.append((1 - t) ** 2 * p0 + 2 * t * (1 - t) * ctrl[0] + t * t * p_end)
            else:
                out.append((1 - t) ** 3 * p0 + 3 * t * (1 - t) ** 2 * ctrl[0] + 3 * t * t * (1 - t) * ctrl[1] + t ** 3 * p_end)
        i += len(ctrl) + 1
    return polygon_area(np.asarray(out))


def flatten(pts, flags, steps=24):
    out, i, n = [], 0, len(pts)
    while i < n:
        if flags[i] == 0:
            out.append(np.asarray(pts[i], np.float64)); i += 1
            continue
        p0, ctrl = out[-1], [np.asarray(pts[i], np.float64)]
        if flags[i + 1] != 0:
            ctrl.append(np.asarray(pts[i + 1], np.float64))
        p_end = np.asarray(pts[i + len(ctrl)], np.float64)
        for s in range(1, steps + 1):
            t = s / steps
            if len(ctrl) == 1:
                out.append((1 - t) ** 2 * p0 + 2 * t * (1 - t) * ctrl[0] + t * t * p_end)
            else:
                out.append((1 - t) ** 3 * p0 + 3 * t * (1 - t) ** 2 * ctrl[0] + 3 * t * t * (1 - t) * ctrl[1] + t ** 3 * p_end)
        i += len(ctrl) + 1
    return np.asarray(out)


def winding_area(pts, flags, offs, lo, hi, step=0.05):
    """Area where the winding number of the (closed) contours is non-zero, by sampling a regular grid."""
    xs = np.arange(lo[0] + step / 2, hi[0], step)
    ys = np.arange(lo[1] + step / 2, hi[1], step)
    gx, gy = np.meshgrid(xs, ys)
    winding = np.zeros(gx.shape, np.int32)
    for c in range(len(offs) - 1):
        poly = flatten(pts[offs[c]:offs[c + 1]], flags[offs[c]:offs[c + 1]])
        a, b = poly, np.roll(poly, -1, axis=0)
        for (x0, y0), (x1, y1) in zip(a, b):
            if y0 == y1:
                continue
            up = (y0 <= gy) & (gy < y1)
            down = (y1 <= gy) & (gy < y0)
            xi = x0 + (gy - y0) * (x1 - x0) / (y1 - y0)
            winding += (up & (xi > gx)).astype(np.int32) - (down & (xi > gx)).astype(np.int32)
    return float((winding != 0).sum()) * step * step


def test_round_caps_are_half_discs():
    """A line with round caps: rectangle plus one disc of the stroke's radius; every arc point lies on the circle
    around the end point (Contour::push_arc_from_unit_chord, outline.rs:632-680)."""
    pts, flags, offs = api.stroke_to_fill([(20, 30), (70, 30)], [0, 0], [0, 2], [0], line_width=8.0, line_cap="round")
    assert len(offs) == 2 and flags.any()
    area = polygon_area_flattened(pts, flags)
    assert abs(area - (50 * 8 + np.pi * 16)) < 0.05
    on_curve = pts[flags == 0]
    right = on_curve[on_curve[:, 0] > 70 + 1e-3]
    left = on_curve[on_curve[:, 0] < 20 - 1e-3]
    assert len(right) and len(left)
    assert np.allclose(np.hypot(right[:, 0] - 70, right[:, 1] - 30), 4.0, atol=1e-4)
    assert np.allclose(np.hypot(left[:, 0] - 20, left[:, 1] - 30), 4.0, atol=1e-4)
    assert abs(pts[:, 0].max() - 74) < 1e-3 and abs(pts[:, 0].min() - 16) < 1e-3


def test_joins_fill_the_outer_corner():
    """A right-angle polyline, filled with the winding rule: the union of the two legs, plus on the outer side of
    the corner the full square (miter), half of it (bevel) or a quarter disc (round)."""
    corner = [(10, 10), (60, 10), (60, 50)]
    w, r = 6.0, 3.0
    union = 50 * w + 40 * w - r * r
    expected = {"miter": union + r * r, "bevel": union + 0.5 * r * r, "round": union + 0.25 * np.pi * r * r}
    for join, want in expected.items():
        pts, flags, offs = api.stroke_to_fill(corner, [0] * 3, [0, 3], [0], line_width=w, line_join=join, miter_limit=10.0)
        assert len(offs) == 2
        got = winding_area(pts, flags, offs, (0, 0), (70, 60))
        assert abs(got - want) < 0.15, (join, got, want)


def test_closed_circle_with_round_joins_is_an_annulus():
    """A circle of four cubics, stroked and filled with the winding rule: an annulus. (Round joins are always drawn
    clockwise, so on one side of the path they go the long way round — inside the stroke, where they change nothing.)"""
    k, R = 0.5522847498 * 40, 40.0
    c = [(R, 0), (R, k), (k, R), (0, R), (-k, R), (-R, k), (-R, 0), (-R, -k), (-k, -R), (0, -R), (k, -R), (R, -k), (R, 0)]
    f = [0, 1, 2] * 4 + [0]
    for join in ("round", "bevel"):
        pts, flags, offs = api.stroke_to_fill(np.asarray(c) + 64, f, [0, 13], [1], line_width=10.0, line_join=join)
        assert len(offs) == 3
        got = winding_area(pts, flags, offs, (10, 10), (118, 118), step=0.25)
        assert abs(got - np.pi * (45 ** 2 - 35 ** 2)) < 8.0, (join, got)


def test_unknown_styles_are_refused():
    style = L.PFStrokeStyle(2.0, 7, 0, 4.0)
    pts = np.zeros((2, 2), np.float32)
    flags, offs, closed = np.zeros(2, np.uint8), np.array([0, 2], np.uint32), np.zeros(1, np.uint8)
    import ctypes as C
    assert not L.lib().PFOutlineStrokeToFill(pts.ctypes.data, flags.ctypes.data, offs.ctypes.data, closed.ctypes.data, 1, C.byref(style))
    assert b"unknown" in L.lib().PFCudaGetLastError()


@pytest.mark.skipif(not os.path.exists(TIGER_SVG), reason="needs the reference's tiger SVG (build container only)")
def test_tiger_strokes_match_the_python_restatement():
    """Every stroked path of the Ghostscript tiger: same contours; the same point flags on all but at most one
    contour (a split decision that sits on the stroker's tolerance falls differently in f32 and in the Python
    restatement's double precision); coordinates within 2e-3 for 99 % of the points and within 0.05 for all (miter
    points of nearly parallel tangents are ill-conditioned)."""
    import sys
    import xml.etree.ElementTree as ET
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_tiger_scene as T
    ns = "{http://www.w3.org/2000/svg}"
    strokes, structural, diffs = 0, 0, []
    for g in ET.parse(TIGER_SVG).getroot().iter(ns + "g"):
        for el in g.findall(ns + "path"):
            stroke = el.attrib.get("stroke")
            if not stroke or stroke == "none":
                continue
            contours = T.parse_path_data(el.attrib["d"])
            width = max(float(el.attrib.get("stroke-width", "1")), T.HAIRLINE_STROKE_WIDTH)
            ref = [o for o in T.stroke_to_fill(contours, width) if o.points]
            points, flags, offsets, closed = [], [], [0], []
            for segs, is_closed in contours:  # Outline::from_segments: the first point, then each segment's tail
                points.append(T.seg_from(segs[0])); flags.append(0)
                for s in segs:
                    if s[0] == "Q":
                        points.append(s[2]); flags.append(1)
                    elif s[0] == "C":
                        points += [s[2], s[3]]; flags += [1, 2]
                    points.append(s[-1]); flags.append(0)
                offsets.append(len(points)); closed.append(1 if is_closed else 0)
            out_p, out_f, out_c = api.stroke_to_fill(points, flags, offsets, closed, width, "miter", 4.0, "butt")
            assert len(out_c) - 1 == len(ref)
            for i, o in enumerate(ref):
                a, b = int(out_c[i]), int(out_c[i + 1])
                if list(out_f[a:b]) != o.flags:
                    assert abs((b - a) - len(o.flags)) <= 3
                    structural += 1
                    continue
                diffs.append(np.abs(out_p[a:b] - np.asarray(o.points)).max(axis=1))
            strokes += 1
    d = np.concatenate(diffs)
    assert strokes == 52 and structural <= 1 and len(d) > 4000
    assert (d < 2e-3).mean() > 0.99 and d.max() < 0.05


def test_stroked_path_in_a_scene_matches_the_oracle_tiling():
    """A stroked polyline pushed into a Scene tiles like the same stroked outline handed to the oracle."""
    from pathfinder_b200.flat_scene import SceneBuilderPy
    from tests import helpers as H
    poly = [(10.5, 12.25), (80.0, 20.0), (60.0, 90.0), (20.0, 70.0)]
    pts, flags, offs = api.stroke_to_fill(poly, [0] * 4, [0, 4], [1], line_width=5.0, line_join="miter", miter_limit=4.0)
    b = SceneBuilderPy((0, 0, 128, 128))
    for i in range(len(offs) - 1):
        c = pts[offs[i]:offs[i + 1]]
        b.move_to(*c[0])
        for p in c[1:]:
            b.line_to(*p)
        b.close()
    b.end_path((10, 20, 30, 255))
    built = H.oracle_build(b.finish("stroked"))
    assert built.alpha_tile_count > 0 and len(built.fills) > 0
    scene = api.Scene()
    scene.set_view_box((0, 0, 128, 128))
    paint = scene.push_paint((10, 20, 30, 255))
    scene.push_stroked_path(poly, [0] * 4, [0, 4], [1], paint, 5.0, "miter", 4.0)
    seen = {}

    def listener(cmd):
        if int(cmd.kind) == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            seen["tiles"] = int(cmd.u.draw_tiles_d3d11.tile_batch_data.tile_count)
            seen["segments"] = int(cmd.u.draw_tiles_d3d11.tile_batch_data.segment_count)

    scene.build(api.BuildOptions(), listener)
    assert seen["segments"] == built.input_segment_count and seen["tiles"] == built.bbox_tile_count


def test_input_the_offsetter_cannot_converge_on_is_refused():
    """Offset::offset (stroke.rs:243-261) recurses until the offset curve is within tolerance or the piece is shorter
    than the tolerance; with coordinates whose squares overflow f32, or that are not numbers, neither ever happens
    (found by fuzzing: the reference would recurse until its stack ran out). The C ABI refuses such input, and gives
    up with an error rather than allocating without bound when finite geometry needs more than a million pieces."""
    curve = [(10, 10), (1e30, 50), (40, 80)]
    for bad in (curve, [(10, 10), (float("nan"), 50), (40, 80)], [(10, 10), (float("inf"), 50), (40, 80)]):
        with pytest.raises(L.PathfinderCudaError):
            api.stroke_to_fill(bad, [0, 1, 0], [0, 3], [0], line_width=3.0)
    with pytest.raises(L.PathfinderCudaError):
        api.stroke_to_fill([(0, 0), (10, 0)], [0, 0], [0, 2], [0], line_width=float("inf"))
    # at 1e6 px an f32 ulp (0.06) is coarser than the tolerance (0.01): the recursion can only end piece by piece
    with pytest.raises(L.PathfinderCudaError) as e:
        api.stroke_to_fill([(0, 0), (5e5, 1e6), (1e6, 0)], [0, 1, 0], [0, 3], [0], line_width=4.0)
    assert "converge" in str(e.value)
    # large geometry that f32 can resolve still strokes
    pts, flags, offs = api.stroke_to_fill([(0, 0), (5e3, 1e4), (1e4, 0)], [0, 1, 0], [0, 3], [0], line_width=4.0)
    assert len(pts) > 8 and np.isfinite(pts).all()
