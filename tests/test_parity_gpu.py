"""GPU parity: the CUDA pipeline (through the C ABI) against the CPU oracle on the same scenes.

Bit-exact: flattened lines, fills (with alpha tile ids in SequentialExecutor order), non-empty tiles
(coords, backdrop, alpha id, path id, colour, ctrl) and the z-buffer. Within tolerance: alpha-tile
coverage and final RGBA (1/255 per channel, BASELINE.json north_star)."""
import numpy as np
import pytest

from pathfinder_b200 import scenes
from pathfinder_b200.flat_scene import FILL_RULE_EVEN_ODD, FILL_RULE_WINDING, SceneBuilderPy
from tests import helpers as H

pytestmark = pytest.mark.gpu

COVERAGE_TOL = 1.0 / 255.0
RGBA_TOL = 1  # in 8-bit units


def check_scene(flat, xf, area_lut, size=None, background=(1.0, 1.0, 1.0, 1.0), check_lines=True):
    built = H.oracle_build(flat, xf, keep_lines=check_lines)
    r, img = H.cuda_render(flat, xf, size=size, background=background)
    w = int(size[0]) if size else int(flat.view_box[2])
    h = int(size[1]) if size else int(flat.view_box[3])

    if check_lines:
        lines, _paths = r.debug_lines()
        ref = np.concatenate([built.path_lines(p) for p in range(flat.n_paths)] or [np.zeros((0, 4), np.float32)])
        # Paths outside the view box are skipped by the D3D11 builder but still flattened by the
        # CPU tiler; compare only when every path was kept.
        if len(lines) == len(ref):
            assert lines.view(np.uint32).tobytes() == ref.view(np.uint32).tobytes(), "flattened lines differ"
    H.assert_records_equal(r.debug_fills(), built.fills, "fills")
    H.assert_records_equal(r.debug_tiles(), built.tiles, "tiles")
    z, rect = r.debug_z_buffer()
    assert rect == built.z_rect
    assert np.array_equal(z, built.z_buffer), "z-buffer differs"
    stats = r.stats()
    assert stats["alpha_tile_count"] == built.alpha_tile_count
    assert stats["fill_count"] == len(built.fills)

    masks = r.debug_alpha_masks()
    ref_masks = built.alpha_masks(area_lut)
    assert masks.shape == ref_masks.shape
    if masks.size:
        assert np.abs(masks - ref_masks).max() <= COVERAGE_TOL, np.abs(masks - ref_masks).max()

    ref_img = built.render(area_lut, w, h, background=background or (0.0, 0.0, 0.0, 0.0))
    diff = np.abs(img.astype(np.int32) - ref_img.astype(np.int32))
    assert diff.max() <= RGBA_TOL, f"max RGBA diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}"
    return r, img, built


def test_single_triangle(area_lut):
    b = SceneBuilderPy((0, 0, 64, 64))
    b.move_to(5.3, 4.1)
    b.line_to(50.7, 20.2)
    b.line_to(20.5, 58.9)
    b.close()
    b.end_path((255, 0, 0, 255))
    check_scene(b.finish("tri"), None, area_lut)


def test_curves_and_overlap(area_lut):
    b = SceneBuilderPy((0, 0, 128, 96))
    b.move_to(10, 10)
    b.cubic_to(120, -20, 140, 120, 20, 80)
    b.quad_to(-30, 40, 10, 10)
    b.close()
    b.end_path((20, 200, 90, 200), FILL_RULE_WINDING)
    b.move_to(64, 5)
    b.line_to(120, 90)
    b.line_to(8, 90)
    b.close()
    b.move_to(64, 30)
    b.line_to(90, 80)
    b.line_to(38, 80)
    b.close()
    b.end_path((0, 0, 255, 255), FILL_RULE_EVEN_ODD)
    b.move_to(0, 0)
    b.line_to(128, 0)
    b.line_to(128, 96)
    b.line_to(0, 96)
    b.close()
    b.end_path((255, 255, 0, 128))
    check_scene(b.finish("curves"), None, area_lut, background=None)


def test_empty_scene(area_lut):
    b = SceneBuilderPy((0, 0, 32, 32))
    flat = b.finish("empty")
    r, img = H.cuda_render(flat, None, background=(0.0, 0.5, 1.0, 1.0))
    assert (img == np.array([0, 128, 255, 255], dtype=np.uint8)).all()


def test_offscreen_and_clipped_paths(area_lut):
    b = SceneBuilderPy((0, 0, 96, 96))
    # crosses every edge of the view box
    b.move_to(-40, 30)
    b.line_to(60, -50)
    b.line_to(150, 48)
    b.line_to(48, 160)
    b.close()
    b.end_path((200, 30, 30, 255))
    # entirely above the view box: only feeds (culled) backdrops
    b.move_to(10, -100)
    b.line_to(80, -100)
    b.line_to(40, -20)
    b.close()
    b.end_path((0, 255, 0, 255))
    # entirely to the left
    b.move_to(-100, 10)
    b.line_to(-20, 40)
    b.line_to(-60, 80)
    b.close()
    b.end_path((0, 0, 255, 255))
    check_scene(b.finish("offscreen"), None, area_lut, check_lines=False)


@pytest.mark.parametrize("size,even_odd", [(256, False), (1024, False), (1024, True)])
def test_tiger(area_lut, size, even_odd):
    flat, xf = scenes.tiger(size, even_odd_odd_paths=even_odd)
    check_scene(flat, xf, area_lut)


@pytest.mark.parametrize("n,size,seed", [(200, 512, 1), (3000, 1024, 0x5EED0004)])
def test_random_paths(area_lut, n, size, seed):
    flat = scenes.random_paths(n, size, seed, r_min=8.0, r_max=96.0)
    check_scene(flat, None, area_lut, background=None)


def test_matches_golden_tiles(area_lut):
    """The committed golden lists (tests/golden/tiger256_*.npy) were produced by the oracle."""
    import os
    g = os.path.join(os.path.dirname(__file__), "golden")
    flat, xf = scenes.tiger(256)
    r, _img = H.cuda_render(flat, xf)
    fills = np.load(os.path.join(g, "tiger256_fills.npy"))
    tiles = np.load(os.path.join(g, "tiger256_tiles.npy"))
    H.assert_records_equal(r.debug_fills(), fills, "fills vs golden")
    H.assert_records_equal(r.debug_tiles(), tiles, "tiles vs golden")
