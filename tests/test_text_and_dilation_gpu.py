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
