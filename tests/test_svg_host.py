"""SVG path data -> outline on the host (front-end half): PFSvgPathDataToOutline against the
Python parser that built the tiger fixture (tools/make_tiger_scene.py) on every path of the tiger, and on synthetic
path data for the grammar corners (relative commands, implicit repetition, smooth curves, arcs)."""
import os
import sys

import numpy as np
import pytest

from pathfinder_b200 import _lib as L
from pathfinder_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIGER_SVG = "/root/reference/resources/svg/Ghostscript_Tiger.svg"
sys.path.insert(0, os.path.join(ROOT, "tools"))


def python_outline(d):
    import make_tiger_scene as T
    points, flags, offsets, closed = [], [], [0], []
    for segs, is_closed in T.parse_path_data(d):
        points.append(T.seg_from(segs[0])); flags.append(0)
        for s in segs:
            if s[0] == "Q":
                points.append(s[2]); flags.append(1)
            elif s[0] == "C":
                points += [s[2], s[3]]; flags += [1, 2]
            points.append(s[-1]); flags.append(0)
        offsets.append(len(points)); closed.append(1 if is_closed else 0)
    return np.asarray(points, np.float32).reshape(-1, 2), np.asarray(flags, np.uint8), offsets, closed


CASES = [
    "M10 10 L20 10 20 20z",
    "m10,10 l10,0 0,10 -10,0z m30 0 h5 v5 h-5z",
    "M0 0C10 0 20 10 20 20S30 40 40 40s5 5 10 10",
    "M0 0Q10 0 10 10T20 20t5-5",
    "M 1e1 -2.5E-1 L.5.5-1-1",
    "M10 10 20 20 30 10",                  # implicit lineto after moveto
    "M0 0 L10 0 Z L5 5 10 5",              # drawing after Z without M
    "M10 50 A30 20 15 0 1 80 60",
    "M10 50 a30 20 -30 1 0 40 10 a5 5 0 0 1 10 0",
    "M10 50 A0 20 0 0 1 80 60",            # zero radius: a line
    "M50 50 A100 100 0 0 1 60 51",         # radii too small for the chord: scaled up
]


@pytest.mark.parametrize("d", CASES)
def test_synthetic_path_data(d):
    pts, flags, offsets, closed = api.svg_path_to_outline(d)
    ref_pts, ref_flags, ref_offsets, ref_closed = python_outline(d)
    assert list(offsets) == ref_offsets and list(closed) == ref_closed and list(flags) == list(ref_flags)
    assert np.abs(pts - ref_pts).max() <= 1e-4 if len(pts) else True


def test_arc_lies_on_its_circle():
    pts, flags, offsets, closed = api.svg_path_to_outline("M60 50 A10 10 0 1 1 50 40")
    assert list(offsets) == [0, len(pts)] and list(closed) == [0]
    on = pts[flags == 0]
    assert np.allclose(np.hypot(on[:, 0] - 50, on[:, 1] - 50), 10.0, atol=1e-4)
    assert (len(pts) - 1) // 3 == 3  # 270 degrees: three cubics


def test_malformed_path_data_is_refused():
    for d in ("M10", "L 10 10 20", "M0 0 A 5 5 0 2 0 10 10", "M0 0 X 5 5", "10 10"):
        with pytest.raises(L.PathfinderCudaError):
            api.svg_path_to_outline(d)
    pts, _f, offsets, _c = api.svg_path_to_outline("")
    assert len(pts) == 0 and list(offsets) == [0]


@pytest.mark.skipif(not os.path.exists(TIGER_SVG), reason="needs the reference's tiger SVG (build container only)")
def test_every_tiger_path_matches_the_python_parser():
    """All 138 path elements: identical structure, coordinates bit-identical (both parse in double precision and
    round to f32 once; the tiger has no arcs, so no libm call is involved)."""
    import xml.etree.ElementTree as ET
    ns = "{http://www.w3.org/2000/svg}"
    n = 0
    for el in ET.parse(TIGER_SVG).getroot().iter(ns + "path"):
        d = el.attrib["d"]
        pts, flags, offsets, closed = api.svg_path_to_outline(d)
        ref_pts, ref_flags, ref_offsets, ref_closed = python_outline(d)
        assert list(offsets) == ref_offsets and list(closed) == ref_closed
        assert np.array_equal(flags, ref_flags) and np.array_equal(pts, ref_pts)
        n += 1
    assert n == 138


@pytest.mark.skipif(not os.path.exists(TIGER_SVG), reason="needs the reference's tiger SVG (build container only)")
def test_tiger_scene_from_the_cpp_front_end_equals_the_fixture():
    """SVG path data -> PFSvgPathDataToOutline -> (PFOutlineStrokeToFill) -> Scene reproduces tests/golden/tiger.npz
    up to the stroker's f32-vs-f64 differences: same paths and colours; fill paths bit-identical."""
    import xml.etree.ElementTree as ET
    import make_tiger_scene as T
    from pathfinder_b200 import scenes
    from pathfinder_b200.flat_scene import FlatScene
    fixture = FlatScene.load(os.path.join(ROOT, "tests", "golden", "tiger.npz"))
    ns = "{http://www.w3.org/2000/svg}"
    path_index = 0
    for g in ET.parse(TIGER_SVG).getroot().iter(ns + "g"):
        inherited = g.attrib.get("fill")
        for el in g.findall(ns + "path"):
            pts, flags, offsets, closed = api.svg_path_to_outline(el.attrib["d"])
            fill, stroke = el.attrib.get("fill", inherited), el.attrib.get("stroke")
            if fill and fill != "none":
                c0, c1 = fixture.path_contour_offsets[path_index], fixture.path_contour_offsets[path_index + 1]
                p0, p1 = fixture.contour_offsets[c0], fixture.contour_offsets[c1]
                assert np.array_equal(fixture.points[p0:p1], pts) and np.array_equal(fixture.point_flags[p0:p1], flags)
                assert tuple(fixture.paint_colors[fixture.paints[path_index]]) == T.parse_color(fill)
                path_index += 1
            if stroke and stroke != "none":
                width = max(float(el.attrib.get("stroke-width", "1")), T.HAIRLINE_STROKE_WIDTH)
                sp, sf, so = api.stroke_to_fill(pts, flags, offsets, closed, width, "miter", 4.0, "butt")
                c0, c1 = fixture.path_contour_offsets[path_index], fixture.path_contour_offsets[path_index + 1]
                p0, p1 = fixture.contour_offsets[c0], fixture.contour_offsets[c1]
                assert abs(int(p1 - p0) - len(sp)) <= 3
                assert tuple(fixture.paint_colors[fixture.paints[path_index]]) == T.parse_color(stroke)
                path_index += 1
    assert path_index == fixture.n_paths == 182


def random_path_data(rng):
    """Grammatical path data: every command with the right number of arguments (sometimes repeated), in the
    number syntaxes SVG allows (signs as separators, leading dots, exponents, commas)."""
    def number(lo=-200.0, hi=200.0):
        v = rng.uniform(lo, hi)
        style = rng.randint(0, 5)
        if style == 0:
            return "%d" % int(v)
        if style == 1:
            return ("%.3f" % v).replace("0.", ".", 1) if abs(v) < 1 else "%.3f" % v
        if style == 2:
            return "%.2e" % v
        return "%g" % round(v, int(rng.randint(0, 4)))

    def sep():
        return str(rng.choice([" ", ",", " , ", "\n", ""]))

    def args(n):
        out = ""
        for i in range(n):
            s = number()
            out += (sep() or (" " if not s.startswith("-") else "")) if i else ""
            if out and out[-1] not in " ,\n" and not s.startswith("-"):
                out += " "
            out += s
        return out

    d = "M" + args(2)
    for _ in range(int(rng.randint(1, 10))):
        c = str(rng.choice(list("MmLlHhVvCcSsQqTtAaZz")))
        n = {"M": 2, "L": 2, "H": 1, "V": 1, "C": 6, "S": 4, "Q": 4, "T": 2, "A": 7, "Z": 0}[c.upper()]
        d += str(rng.choice(["", " "])) + c
        for _rep in range(int(rng.randint(1, 3)) if n else 0):
            if c.upper() == "A":
                d += " " + " ".join([number(1, 80), number(1, 80), number(-180, 180), str(rng.randint(0, 2)),
                                     str(rng.randint(0, 2)), number(), number()])
            else:
                d += " " + args(n)
    return d


def test_fuzzed_path_data_matches_the_python_parser():
    """Differential fuzz of the two independent parsers (C++ in csrc/svg.cpp, Python in tools/make_tiger_scene.py):
    same contours, flags and closedness; coordinates equal up to the libm calls of the arc conversion."""
    rng = np.random.RandomState(21)
    compared = 0
    for _ in range(400):
        d = random_path_data(rng)
        pts, flags, offsets, closed = api.svg_path_to_outline(d)
        ref_pts, ref_flags, ref_offsets, ref_closed = python_outline(d)
        assert list(offsets) == ref_offsets, d
        assert list(closed) == ref_closed and list(flags) == list(ref_flags), d
        if len(pts):
            scale = max(1.0, float(np.abs(ref_pts).max()))
            assert np.abs(pts - ref_pts).max() <= 1e-5 * scale, d
            compared += len(pts)
    assert compared > 5000
