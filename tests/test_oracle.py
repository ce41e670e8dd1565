"""CPU oracle (oracle/pf_oracle.cpp) checks. The reference holds no golden vectors for this path
(parity unpinned), so the oracle is pinned by (i) the reference's simd known answers
(simd/src/test.rs:54-60,380-383: floor / ceil / round-to-nearest-even conversions) observed through
the tiler, (ii) geometric properties, and (iii) committed golden lists that fix its output."""
import os

import numpy as np
import pytest

from pathfinder_b200 import scenes
from pathfinder_b200.flat_scene import SceneBuilderPy
from tests import helpers as H

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
INVALID = 0xFFFFFFFF


def polygon_scene(points, size=64, rgba=(255, 255, 255, 255), fill_rule=0):
    b = SceneBuilderPy((0, 0, size, size))
    b.move_to(*points[0])
    for p in points[1:]:
        b.line_to(*p)
    b.close()
    b.end_path(rgba, fill_rule)
    return b.finish("poly")


def shoelace(points):
    a = 0.0
    for (x0, y0), (x1, y1) in zip(points, points[1:] + points[:1]):
        a += x0 * y1 - x1 * y0
    return abs(a) / 2.0


def test_golden_lists_are_stable():
    flat, xf = scenes.tiger(256)
    b = H.oracle_build(flat, xf)
    H.assert_records_equal(b.fills, np.load(os.path.join(GOLDEN, "tiger256_fills.npy")), "fills")
    H.assert_records_equal(b.tiles, np.load(os.path.join(GOLDEN, "tiger256_tiles.npy")), "tiles")
    assert np.array_equal(b.z_buffer, np.load(os.path.join(GOLDEN, "tiger256_zbuffer.npy")))


def test_fill_quantisation_rounds_half_to_even():
    """_mm_cvtps_epi32 semantics (simd/src/x86/mod.rs:318-320; KATs simd/src/test.rs:380-383):
    x.5 rounds to the even integer. Vertices are placed at tile-relative k + 0.5 (in 1/256 px)."""
    f = np.float32
    rel = [(2.5, 10.5), (200.5, 77.5), (101.5, 3000.5)]  # 1/256 px inside tile (1, 1)
    pts = [(16 + x / 256.0, 16 + y / 256.0) for x, y in rel]
    b = H.oracle_build(polygon_scene(pts), None)
    expect = []
    for (x0, y0), (x1, y1) in zip(pts, pts[1:] + pts[:1]):
        q = [int(np.rint((f(v) - f(16.0)) * f(256.0))) for v in (x0, y0, x1, y1)]
        expect.append(tuple(q))
    got = [(int(r["from_x"]), int(r["from_y"]), int(r["to_x"]), int(r["to_y"])) for r in b.fills]
    assert got == expect
    assert (2, 10, 200, 78) == expect[0]  # 2.5 -> 2, 10.5 -> 10, 200.5 -> 200, 77.5 -> 78
    assert len(b.tiles) == 1 and b.tiles[0]["tile_x"] == 1 and b.tiles[0]["tile_y"] == 1
    assert b.alpha_tile_count == 1 and (b.fills["link"] == 0).all()


def test_tile_rect_floor_ceil():
    """round_out uses floor/ceil (simd/src/test.rs:54-60): a bound exactly on a tile edge does not
    spill into the next tile."""
    b = H.oracle_build(polygon_scene([(17.0, 5.0), (48.0, 5.0), (48.0, 31.9), (17.0, 31.9)]), None)
    assert b.bbox_tile_count == 2 * 2
    b = H.oracle_build(polygon_scene([(17.0, 5.0), (48.1, 5.0), (48.1, 32.0), (17.0, 32.0)]), None)
    assert b.bbox_tile_count == 3 * 2


def test_rectangle_tiles_and_coverage(area_lut):
    rect = [(8.0, 8.0), (56.0, 8.0), (56.0, 40.0), (8.0, 40.0)]
    b = H.oracle_build(polygon_scene(rect), None)
    tiles = {(int(t["tile_x"]), int(t["tile_y"])): t for t in b.tiles}
    assert set(tiles) == {(x, y) for x in range(4) for y in range(3)}
    for (x, y), t in tiles.items():
        interior = x in (1, 2) and y == 1
        assert (t["alpha_tile_id"] == INVALID) == interior
        if interior:
            assert abs(int(t["backdrop"])) == 1  # solid tile: winding number of the interior
    img = b.render(area_lut, 64, 64)
    alpha = img[:, :, 3].astype(np.int32)
    ideal = np.zeros((64, 64), dtype=np.int32)
    ideal[8:40, 8:56] = 255
    # The LUT is sampled 1/32 px off-centre (utils/area-lut/src/main.rs:84-85), so edges aligned to
    # pixel boundaries are reproduced within a few 1/255 steps, not exactly.
    assert np.abs(alpha - ideal).max() <= 10
    assert (alpha[12:36, 12:52] >= 253).all() and (alpha[44:, :] <= 2).all()  # 8-bit LUT: sums land within 2 steps


@pytest.mark.parametrize("seed", range(6))
def test_coverage_integrates_to_polygon_area(area_lut, seed):
    """Sum of per-pixel coverage = polygon area (the point of exact-area coverage)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 9))
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = rng.uniform(20, 58, n)
    pts = [(64 + float(r * np.cos(a)), 64 + float(r * np.sin(a))) for a, r in zip(ang, rad)]
    b = H.oracle_build(polygon_scene(pts, size=128), None)
    img = b.render(area_lut, 128, 128)
    area = img[:, :, 3].astype(np.float64).sum() / 255.0
    assert abs(area - shoelace(pts)) <= 0.01 * shoelace(pts) + 2.0


def test_even_odd_and_winding_differ_on_self_overlap(area_lut):
    star = [(64, 8), (98, 112), (10, 44), (118, 44), (30, 112)]
    w = H.oracle_build(polygon_scene(star, size=128, fill_rule=0), None).render(area_lut, 128, 128)
    e = H.oracle_build(polygon_scene(star, size=128, fill_rule=1), None).render(area_lut, 128, 128)
    assert w[60, 64, 3] == 255 and e[60, 64, 3] == 0  # centre pentagon has winding number 2
    assert w[30, 64, 3] == 255 and e[30, 64, 3] == 255  # a point of the star


def test_flattening_properties():
    b = SceneBuilderPy((0, 0, 512, 512))
    b.move_to(20, 20)
    b.cubic_to(500, -100, 520, 600, 40, 480)
    b.quad_to(-100, 250, 20, 20)
    b.close()
    b.end_path((0, 0, 0, 255))
    built = H.oracle_build(b.finish(), None, keep_lines=True)
    lines = built.path_lines(0)
    assert built.line_segment_count == len(lines) > 20
    # consecutive lines join up; the contour closes with a zero-length line
    assert np.array_equal(lines[1:, :2], lines[:-1, 2:])
    assert tuple(lines[0, :2]) == (20.0, 20.0) and tuple(lines[-1, 2:]) == (20.0, 20.0)
    # every chord is short enough to satisfy the flatness tolerance scale (0.25 px deviation)
    assert np.hypot(lines[:, 2] - lines[:, 0], lines[:, 3] - lines[:, 1]).max() < 80.0


def test_off_screen_and_degenerate_inputs():
    b = SceneBuilderPy((0, 0, 64, 64))
    b.end_path((1, 2, 3, 255))           # empty path
    b.move_to(5, 5)
    b.close()
    b.end_path((1, 2, 3, 255))           # single point
    b.move_to(-50, -50)
    b.line_to(-10, -50)
    b.line_to(-30, -10)
    b.close()
    b.end_path((1, 2, 3, 255))           # entirely off screen
    built = H.oracle_build(b.finish(), None)
    assert len(built.fills) == 0 and len(built.tiles) == 0 and built.alpha_tile_count == 0


def test_zero_area_path_has_fills_but_no_coverage(area_lut):
    """add_fill culls only on from_x == to_x (renderer/src/builder.rs:532-536): a degenerate
    horizontal path still emits fills, which cancel in the mask."""
    built = H.oracle_build(polygon_scene([(5.0, 5.0), (50.0, 5.0)]), None)
    assert len(built.fills) > 0 and built.alpha_tile_count > 0
    assert built.render(area_lut, 64, 64)[:, :, 3].max() <= 1


def test_threaded_build_matches_sequential_up_to_alpha_ids():
    """RayonExecutor stand-in: alpha tile ids are racy (gpu_data.rs:449-454) but everything else
    is identical after renumbering ids in first-use order."""
    flat, xf = scenes.tiger(512)
    seq = H.oracle_build(flat, xf)
    par = H.oracle_build(flat, xf, n_threads=4)
    assert len(seq.fills) == len(par.fills) and seq.alpha_tile_count == par.alpha_tile_count

    def canon(fills, tiles):
        remap = {}
        links = np.empty(len(fills), dtype=np.uint32)
        for i, l in enumerate(fills["link"].tolist()):
            links[i] = remap.setdefault(l, len(remap))
        f = fills.copy()
        f["link"] = links
        t = tiles.copy()
        t["alpha_tile_id"] = [remap[a] if a != INVALID else INVALID for a in tiles["alpha_tile_id"].tolist()]
        return f, t

    fs, ts = canon(seq.fills, seq.tiles)
    fp, tp = canon(par.fills, par.tiles)
    H.assert_records_equal(fs, fp, "fills")
    H.assert_records_equal(ts, tp, "tiles")
    H.assert_records_equal(fs, seq.fills, "sequential ids are already canonical")


def test_strip_restriction_partitions_the_tile_lists(area_lut):
    """What each rank of the strip partition must produce: the union over strips of (tile coords,
    backdrop, path, solid/alpha) equals the full build, and stitched strips equal the full frame."""
    flat, xf = scenes.tiger(256)
    full = H.oracle_build(flat, xf)
    full_img = full.render(area_lut, 256, 256, background=(1, 1, 1, 1))
    key = lambda t: (int(t["path_id"]), int(t["tile_y"]), int(t["tile_x"]), int(t["backdrop"]), int(t["alpha_tile_id"] == INVALID))
    want = sorted(key(t) for t in full.tiles)
    got, fills = [], 0
    stitched = np.zeros_like(full_img)
    for y0, y1 in [(0, 4), (4, 8), (8, 12), (12, 16)]:
        part = H.oracle_build(flat, xf, strip=(y0, y1))
        assert all(y0 <= int(t["tile_y"]) < y1 for t in part.tiles)
        got += [key(t) for t in part.tiles]
        fills += len(part.fills)
        img = part.render(area_lut, 256, 256, background=(1, 1, 1, 1))
        stitched[y0 * 16:y1 * 16] = img[y0 * 16:y1 * 16]
    assert sorted(got) == want
    assert fills == len(full.fills)
    assert np.array_equal(stitched, full_img)


def test_golden_tools_round_trip(tmp_path):
    """tools/make_golden_lists.py regenerates the committed golden lists exactly, and the text dump used
    to pin the oracle against the Rust tiler (tools/dump_lists.py + tools/diff_lists.py) round-trips."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    run = lambda *a: subprocess.run([sys.executable, *a], cwd=root, capture_output=True, text=True)
    r = run("tools/make_golden_lists.py", "--check")
    assert r.returncode == 0, r.stdout + r.stderr
    a, b, scene = str(tmp_path / "a.lists"), str(tmp_path / "b.lists"), str(tmp_path / "a.scene")
    assert run("tools/dump_lists.py", "tiger", "256", "--out", a, "--scene-out", scene).returncode == 0
    assert run("tools/dump_lists.py", "tiger", "256", "--out", b).returncode == 0
    r = run("tools/diff_lists.py", a, b)
    assert r.returncode == 0 and "IDENTICAL" in r.stdout
    # a corrupted record must be reported
    lines = open(b).read().splitlines()
    w = lines[0].split()
    w[1] = str(int(w[1]) ^ 1)
    lines[0] = " ".join(w)
    open(b, "w").write("\n".join(lines) + "\n")
    assert run("tools/diff_lists.py", a, b).returncode == 1
    assert open(scene).readline().startswith("viewbox ")
    # clipped scenes: clip paths in the scene file, Clip records in the lists
    c, cs = str(tmp_path / "c.lists"), str(tmp_path / "c.scene")
    assert run("tools/dump_lists.py", "clips", "256", "--out", c, "--scene-out", cs).returncode == 0
    text = open(c).read()
    assert "\nclip " in text and "clippath " in open(cs).read()
    assert run("tools/diff_lists.py", c, c).returncode == 0
    broken = str(tmp_path / "c2.lists")
    open(broken, "w").write(text.replace("\nclip ", "\nclip 1", 1))
    assert run("tools/diff_lists.py", c, broken).returncode == 1


def _clip_scene(clip_kind):
    b = SceneBuilderPy((0, 0, 128, 128))
    if clip_kind == "cover":      # covers the whole view box: clipping changes nothing
        b.move_to(-10, -10); b.line_to(138, -10); b.line_to(138, 138); b.line_to(-10, 138); b.close()
    elif clip_kind == "miss":     # lies outside every draw tile: everything is clipped away
        b.move_to(100, 100); b.line_to(126, 100); b.line_to(126, 126); b.line_to(100, 126); b.close()
    else:                         # a diamond through the middle of the draw paths
        b.move_to(64, 6); b.line_to(122, 64); b.line_to(64, 122); b.line_to(6, 64); b.close()
    clip = b.end_clip_path()
    b.move_to(8, 8); b.line_to(90, 12); b.line_to(80, 88); b.line_to(12, 70); b.close()
    b.end_path((200, 30, 30, 255), clip=clip)
    b.move_to(20, 20); b.cubic_to(90, 0, 100, 90, 30, 80); b.close()
    b.end_path((30, 30, 200, 160), clip=clip)
    return b.finish(clip_kind)


def test_clip_path_cases(area_lut):
    """Tiler::prepare_tiles clip cases (renderer/src/tiler.rs:114-156) + D3D9 clip combine, through properties:
    a clip path covering everything leaves the frame unchanged, one that misses every draw tile erases it, and
    a partial clip equals the unclipped frame inside the clip path and the background outside, away from its edge."""
    import dataclasses
    bg = (1.0, 1.0, 1.0, 1.0)
    flat = _clip_scene("cover")
    unclipped = dataclasses.replace(flat, draw_clip_paths=np.full(flat.n_paths, 0xFFFFFFFF, np.uint32))
    ref = H.oracle_build(unclipped).render(area_lut, 128, 128, background=bg)
    covered = H.oracle_build(flat)
    assert np.array_equal(covered.render(area_lut, 128, 128, background=bg), ref)
    assert len(covered.clips) == 0  # the clip path has no alpha tile under the draw paths

    missed = H.oracle_build(_clip_scene("miss"))
    assert len(missed.tiles) == 0
    assert (missed.render(area_lut, 128, 128, background=bg) == 255).all()

    part = H.oracle_build(_clip_scene("diamond"))
    assert len(part.clips) > 0
    img = part.render(area_lut, 128, 128, background=bg).astype(np.int32)
    yy, xx = np.mgrid[0:128, 0:128]
    d = np.abs(xx + 0.5 - 64) + np.abs(yy + 0.5 - 64)   # L1 distance from the diamond's centre (radius 58)
    inside, outside = d < 55, d > 61
    assert np.abs(img[inside] - ref.astype(np.int32)[inside]).max() <= 1
    assert (img[outside] >= 254).all()  # (the LUT's 1/32 px bias leaves 1 LSB at a few tile-edge columns)
    # every Clip record joins two alpha tiles that exist, and the combined tiles carry a zero backdrop
    assert part.clips["dest_tile_id"].max() < part.alpha_tile_count and part.clips["src_tile_id"].max() < part.alpha_tile_count
    combined = part.tiles[np.isin(part.tiles["alpha_tile_id"], part.clips["dest_tile_id"])]
    assert (combined["backdrop"] == 0).all()


@pytest.mark.skipif(not os.path.exists("/root/reference/resources/svg/Ghostscript_Tiger.svg"),
                    reason="the reference checkout (and its tiger SVG) is only present in the build container")
def test_tiger_fixture_is_reproducible():
    """tests/golden/tiger.npz is exactly what tools/make_tiger_scene.py derives from the reference's
    resources/svg/Ghostscript_Tiger.svg (path-data parser + stroke-to-fill restatement)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "tools/make_tiger_scene.py", "--check"], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_strip_partition_and_threading_on_fuzzed_scenes(area_lut):
    """The oracle itself, on the seeded fuzz scenes the GPU tests use: uneven strips partition the tile lists and
    reproduce the frame; a threaded build equals the sequential one up to alpha tile numbering."""
    from tests.test_parity_gpu import fuzz_scene
    key = lambda t: (int(t["path_id"]), int(t["tile_y"]), int(t["tile_x"]), int(t["backdrop"]), int(t["alpha_tile_id"] == INVALID))
    for seed in range(0, 60, 3):
        flat, xf, (w, h) = fuzz_scene(seed, rotate=seed % 2 == 0)
        full = H.oracle_build(flat, xf)
        full_img = full.render(area_lut, w, h, background=(1, 1, 1, 1))
        rows = (h + 15) // 16
        cuts = sorted({0, rows // 3, (2 * rows + 2) // 3, rows})
        got, fills = [], 0
        stitched = np.zeros_like(full_img)
        for y0, y1 in zip(cuts[:-1], cuts[1:]):
            part = H.oracle_build(flat, xf, strip=(y0, y1))
            got += [key(t) for t in part.tiles]
            fills += len(part.fills)
            img = part.render(area_lut, w, h, background=(1, 1, 1, 1))
            stitched[y0 * 16:y1 * 16] = img[y0 * 16:y1 * 16]
        assert sorted(got) == sorted(key(t) for t in full.tiles), seed
        assert fills == len(full.fills) and np.array_equal(stitched, full_img), seed
        threaded = H.oracle_build(flat, xf, n_threads=4)
        assert sorted(key(t) for t in threaded.tiles) == sorted(key(t) for t in full.tiles), seed
        assert len(threaded.fills) == len(full.fills) and threaded.alpha_tile_count == full.alpha_tile_count, seed


@pytest.mark.skipif(not os.path.exists("/root/reference/resources/fonts/Roboto-Regular.ttf"),
                    reason="the reference checkout (and its fonts) is only present in the build container")
def test_glyph_fixture_is_reproducible():
    """tests/golden/roboto_glyphs.npz is what tools/make_glyph_fixture.py reads out of the reference's Roboto-Regular.ttf."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "tools/make_glyph_fixture.py", "--check"], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def _contour_area(pts, flags):
    """Signed area of a closed contour of lines and quadratics (flag 1 = control point): Green's theorem, with
    integral(B x B') = 2/3 (p0 x c) + 2/3 (c x p1) + 1/3 (p0 x p1) for a quadratic."""
    cross = lambda a, b: float(a[0]) * float(b[1]) - float(a[1]) * float(b[0])
    total, i, n = 0.0, 0, len(pts)
    while i + 1 < n:
        if flags[i + 1] == 1:
            p0, c, p1 = pts[i], pts[i + 1], pts[i + 2]
            total += (2 * cross(p0, c) + 2 * cross(c, p1) + cross(p0, p1)) / 3.0
            i += 2
        else:
            total += cross(pts[i], pts[i + 1])
            i += 1
    total += cross(pts[n - 1], pts[0])  # the implicit closing line
    return total / 2.0


@pytest.mark.parametrize("layout", ["grid", "lines"])
def test_text_page_scene(area_lut, layout):
    """The text-page scene (BASELINE.json configs[2], outlines only): one small winding-rule path per glyph. The
    rendered ink equals the analytic area of the glyph outlines (holes are opposite-wound contours of the same path)."""
    n, size = 300, 384
    flat = scenes.text_page(n, size, layout=layout)
    assert len(flat.fill_rules) == n and (flat.fill_rules == 0).all()
    again = scenes.text_page(n, size, layout=layout)
    assert np.array_equal(flat.points, again.points) and np.array_equal(flat.contour_offsets, again.contour_offsets)
    assert flat.points.min() >= 0 and flat.points.max() <= size
    assert (np.asarray(flat.point_flags) <= 1).all()  # TrueType outlines: lines and quadratics only

    co = np.asarray(flat.contour_offsets)
    ink = 0.0
    for p in range(n):
        c0, c1 = int(flat.path_contour_offsets[p]), int(flat.path_contour_offsets[p + 1])
        ink += abs(sum(_contour_area(flat.points[co[c]:co[c + 1]], flat.point_flags[co[c]:co[c + 1]]) for c in range(c0, c1)))

    b = H.oracle_build(flat, None)
    assert len(b.fills) > 0 and b.alpha_tile_count > 0
    img = b.render(area_lut, size, size)
    assert (img[:, :, :3][img[:, :, 3] > 0] == 0).all()  # black ink only
    rendered = img[:, :, 3].astype(np.float64).sum() / 255.0
    assert abs(rendered - ink) <= 0.02 * ink
