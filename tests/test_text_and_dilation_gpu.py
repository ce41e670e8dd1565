"""GPU parity for the text-page scene (BASELINE.json configs[2], outlines) and host-side dilation (stem
darkening): more small paths, or already-prepared points under an identity transform, through the same device
pipeline. Their host halves are pinned on the CPU (tests/test_oracle.py, tests/test_dilate_host.py). First seen
green on a B200 at the end of round 1 (as XPASS) and again at the start of round 2."""
import numpy as np
import pytest

from pathfinder_b200 import scenes
from tests import helpers as H
from tests.test_parity_gpu import COVERAGE_TOL, RGBA_TOL

pytestmark = pytest.mark.gpu


def check(flat, xf, area_lut, dilation=(0.0, 0.0)):
    w, h = int(flat.view_box[2]), int(flat.view_box[3])
    background = (1.0, 1.0, 1.0, 1.0)
    built = H.oracle_build(flat, xf, keep_lines=True, dilation=dilation)
    r, img = H.cuda_render(flat, xf, background=background, dilation=dilation)
    H.assert_lines_match(r, built, flat.n_paths)
    H.assert_records_equal(r.debug_fills(), built.fills, "fills")
    H.assert_records_equal(r.debug_tiles(), built.tiles, "tiles")
    z, rect = r.debug_z_buffer()
    assert rect == built.z_rect and np.array_equal(z, built.z_buffer)
    masks, ref_masks = r.debug_alpha_masks(), built.alpha_masks(area_lut)
    assert masks.shape == ref_masks.shape and np.abs(masks - ref_masks).max() <= COVERAGE_TOL
    ref_img = built.render(area_lut, w, h, background=background)
    assert np.abs(img.astype(np.int32) - ref_img.astype(np.int32)).max() <= RGBA_TOL
    r2, img2 = H.cuda_render(flat, xf, background=background, debug=False, dilation=dilation)
    assert np.array_equal(img2, img), "production path differs from the instrumented path"
    r.close()
    r2.close()


@pytest.mark.parametrize("layout", ["grid", "lines"])
def test_text_page(area_lut, layout):
    check(scenes.text_page(2000, 1024, layout=layout), None, area_lut)


def test_text_page_full_size(area_lut):
    """BASELINE.json configs[2] at its full size (outlines only): 10,000 glyphs on 2048 x 2048."""
    check(scenes.text_page(10000, 2048), None, area_lut)


def test_dilated_tiger(area_lut):
    flat, xf = scenes.tiger(512)
    check(flat, xf, area_lut, dilation=(0.35, 0.2))


def test_stem_darkened_text(area_lut):
    """Stem darkening as the reference's demo sets it: dilation = STEM_DARKENING_FACTORS * font size
    (content/src/effects.rs:32; demo/common/src/lib.rs:273-276), here for 16 px."""
    check(scenes.text_page(2000, 1024, layout="lines"), None, area_lut, dilation=(0.0121 * 16, 0.0121 * 1.25 * 16))


def test_svg_paths_and_strokes_render(area_lut):
    """SURVEY.md §8 f2 end to end on the device: SVG path data (relative commands, smooth curves, arcs) through
    PFSvgPathDataToOutline, strokes of every join and cap through PFOutlineStrokeToFill, into a Scene, through the
    CUDA pipeline — lists bit-exact and RGBA within 1/255 of the oracle given the same outlines."""
    from pathfinder_b200 import api
    from pathfinder_b200.flat_scene import FILL_RULE_EVEN_ODD, FILL_RULE_WINDING, SceneBuilderPy
    shapes = [
        ("M40 40 h120 v90 h-120z M70 60 v50 h60 v-50z", (200, 40, 40, 255), FILL_RULE_EVEN_ODD, None),
        ("M30 200 C80 120 160 280 220 190 S300 120 330 210 q10 40 -40 50 t-60 -20z", (30, 90, 200, 200), FILL_RULE_WINDING, None),
        ("M260 60 a50 30 20 1 0 80 20 a20 20 0 0 1 -80 -20z", (20, 160, 60, 230), FILL_RULE_WINDING, None),
        ("M40 320 L120 260 200 330 280 250 360 330", (0, 0, 0, 255), FILL_RULE_WINDING, (9.0, "miter", "butt")),
        ("M60 360 C120 300 180 420 240 350", (160, 0, 160, 180), FILL_RULE_WINDING, (14.0, "round", "round")),
        ("M300 280 l40 60 l30 -70", (220, 120, 0, 255), FILL_RULE_WINDING, (6.0, "bevel", "square")),
        ("M200 100 a40 40 0 1 1 0.1 0z", (90, 90, 90, 128), FILL_RULE_WINDING, (3.5, "miter", "butt")),
    ]
    b = SceneBuilderPy((0, 0, 704, 704))
    for d, rgba, rule, stroke in shapes:
        pts, flags, offsets, closed = api.svg_path_to_outline(d)
        if stroke is not None:
            width, join, cap = stroke
            pts, flags, offsets = api.stroke_to_fill(pts, flags, offsets, closed, width, line_join=join, line_cap=cap)
        b.add_outline(pts, flags, offsets)
        b.end_path(rgba, rule)
    flat = b.finish("svg-strokes")
    assert flat.n_paths == len(shapes)
    check(flat, (1.7, 0.0, 0.0, 1.7, 10.0, -20.0), area_lut)
