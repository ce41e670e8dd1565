"""Double-entry check of the oracle: tests/py_tiler.py (a second restatement of the CPU tiler, written straight from
the Rust sources in scalar numpy float32) and oracle/pf_oracle.cpp must produce the same flattened lines, the same
fills in the same order with the same alpha-tile ids, and the same backdrops, bit for bit."""
import numpy as np
import pytest

from pathfinder_b200 import scenes
from pathfinder_b200.flat_scene import FlatScene, SceneBuilderPy
from tests import helpers as H
from tests import py_tiler as P

INVALID = 0xFFFFFFFF


def compare(flat: FlatScene, oracle_input=None):
    """`oracle_input` = (scene, transform) when the oracle is to apply a BuildOptions transform itself; `flat` then
    holds the same points already transformed for the second tiler, which has no transform stage."""
    built = H.oracle_build(*(oracle_input or (flat, None)), keep_lines=True)
    mine = P.tile_scene(flat)
    for p in range(flat.n_clip_paths + flat.n_paths):   # clip paths first
        want = built.path_lines(p)
        got = np.asarray(mine["lines"][p], np.float32).reshape(-1, 4)
        assert want.tobytes() == got.tobytes(), f"path {p}: flattened lines differ"
    fills = mine["fills"]
    assert len(fills) == len(built.fills)
    got = np.asarray(fills, np.int64).reshape(-1, 5)
    for k, name in enumerate(("from_x", "from_y", "to_x", "to_y", "link")):
        assert np.array_equal(got[:, k], built.fills[name].astype(np.int64)), name
    # the whole D3D9 batch: every non-empty tile in path order, the Clip records, the z-buffer
    got = np.asarray(mine["tiles"], np.int64).reshape(-1, 5)
    assert len(got) == len(built.tiles)
    for k, name in enumerate(("tile_x", "tile_y", "alpha_tile_id", "path_id", "backdrop")):
        assert np.array_equal(got[:, k], built.tiles[name].astype(np.int64)), name
    paints, _ = flat.palette()
    assert np.array_equal(built.tiles["color"], np.asarray(paints)[built.tiles["path_id"]])
    assert np.array_equal(built.tiles["ctrl"], np.where(flat.fill_rules[built.tiles["path_id"]] == 1, 2, 1))
    got = np.asarray(mine["clips"], np.int64).reshape(-1, 4)
    assert len(got) == len(built.clips)
    for k, name in enumerate(("dest_tile_id", "dest_backdrop", "src_tile_id", "src_backdrop")):
        assert np.array_equal(got[:, k], built.clips[name].astype(np.int64)), name
    x0, y0, x1, y1 = mine["z_rect"]
    assert tuple(built.z_rect) == (x0, y0, x1, y1) or tuple(built.z_rect) == (x0, y0, x1 - x0, y1 - y0)
    assert np.array_equal(mine["z_buffer"], built.z_buffer)
    return len(fills)


def random_scene(seed):
    rng = np.random.RandomState(seed)
    w, h = int(rng.choice([48, 64, 100, 130])), int(rng.choice([48, 64, 90]))
    b = SceneBuilderPy((0.0, 0.0, w, h))
    for _ in range(int(rng.randint(1, 6))):
        for _ in range(int(rng.randint(1, 3))):
            n = int(rng.randint(2, 7))
            spread = float(rng.choice([0.6, 1.0, 1.8]))     # some contours leave the view box on every side
            pt = lambda: ((rng.uniform(-0.5, 0.5) * spread + 0.5) * w, (rng.uniform(-0.5, 0.5) * spread + 0.5) * h)
            snap = rng.rand() < 0.3                          # points on tile corners and pixel centres
            q = (lambda v: (round(v[0] / 8) * 8.0, round(v[1] / 8) * 8.0)) if snap else (lambda v: v)
            b.move_to(*q(pt()))
            for _ in range(n):
                kind = rng.randint(0, 3)
                if kind == 0:
                    b.line_to(*q(pt()))
                elif kind == 1:
                    b.quad_to(*pt(), *q(pt()))
                else:
                    b.cubic_to(*pt(), *pt(), *q(pt()))
            b.close()
        alpha = 255 if rng.rand() < 0.5 else int(rng.randint(40, 250))
        b.end_path((int(rng.randint(0, 256)), int(rng.randint(0, 256)), int(rng.randint(0, 256)), alpha), int(rng.randint(0, 2)))
    return b.finish(f"py{seed}")


@pytest.mark.parametrize("block", range(4))
def test_random_scenes(block):
    total = 0
    for seed in range(block * 12, block * 12 + 12):
        total += compare(random_scene(seed))
    assert total > 500


def test_axis_aligned_and_degenerate_edges():
    b = SceneBuilderPy((0, 0, 64, 64))
    b.move_to(16, 16); b.line_to(48, 16); b.line_to(48, 48); b.line_to(16, 48); b.close()      # on tile boundaries
    b.end_path((255, 0, 0, 255))
    b.move_to(8, 8); b.line_to(8, 8); b.line_to(40, 8); b.line_to(40, 8); b.close()              # zero-area, repeats
    b.end_path((0, 255, 0, 255))
    b.move_to(-20, 32); b.line_to(90, 32.5); b.line_to(90, -5); b.line_to(-20, -5); b.close()    # crosses every edge
    b.end_path((0, 0, 255, 128), 1)
    b.move_to(100, 100); b.line_to(120, 100); b.line_to(110, 130); b.close()                     # outside the view box
    b.end_path((9, 9, 9, 255))
    assert compare(b.finish("edges")) > 0


def test_clipped_scenes():
    """The four cases of Tiler::prepare_tiles with a clip path, Clip records included, on the seeded clip scenes of
    the GPU fuzz (clip paths used in scene order, draw paths clipped or not) with their transforms dropped."""
    from tests.test_parity_gpu import clip_scene, fuzz_clip_scene
    assert compare(clip_scene(128)) > 0
    n_clips = 0
    for seed in range(12):
        flat = fuzz_clip_scene(seed)
        flat = flat[0] if isinstance(flat, tuple) else flat
        compare(flat)
        n_clips += len(H.oracle_build(flat, None).clips)
    assert n_clips > 20


def test_tiger_and_text_page():
    """The tiger at 128 px (transform applied to the points beforehand, in f32 with Transform2F's operation order)
    and a small text page: cubics and quadratics at realistic scales."""
    from tests.test_dilate_host import prepared_scene
    flat, xf = scenes.tiger(128)
    assert compare(prepared_scene(flat, xf, (0.0, 0.0))) > 1000
    # and with the oracle applying the transform itself (Scene::apply_render_options): Transform2F * Vector2F is
    # (m11 x + m12 y) + tx, (m21 x + m22 y) + ty, one rounding per operation
    assert compare(prepared_scene(flat, xf, (0.0, 0.0)), oracle_input=(flat, xf)) > 1000
    rotated = (0.8, -0.45, 0.3, 0.9, 20.0, 5.0)
    text = scenes.text_page(40, 128, layout="lines")
    assert compare(prepared_scene(text, rotated, (0.0, 0.0)), oracle_input=(text, rotated)) > 300
    assert compare(scenes.text_page(60, 128, layout="lines")) > 500


def compare_pixels(flat, area_lut, background=(1.0, 1.0, 1.0, 1.0)):
    w, h = int(flat.view_box[2]), int(flat.view_box[3])
    built = H.oracle_build(flat, None)
    mine = P.tile_scene(flat)
    masks = P.alpha_masks(mine["fills"], built.alpha_tile_count, area_lut)
    want = built.alpha_masks(area_lut)
    assert masks.shape == want.shape
    assert np.abs(masks - want).max() <= 2e-5, np.abs(masks - want).max()
    frame = P.render(flat, mine, area_lut, w, h, background)
    ref = built.render(area_lut, w, h, background=background, want_f32=True)
    ref = ref[1] if isinstance(ref, tuple) else ref
    assert np.abs(frame - ref.reshape(h, w, 4)).max() <= 1e-4


@pytest.mark.parametrize("seed", range(8))
def test_coverage_and_composite_random(area_lut, seed):
    compare_pixels(random_scene(100 + seed), area_lut)


def test_coverage_and_composite_clipped(area_lut):
    from tests.test_parity_gpu import clip_scene, fuzz_clip_scene
    compare_pixels(clip_scene(128), area_lut)
    for seed in range(4):
        flat = fuzz_clip_scene(seed)
        compare_pixels(flat[0] if isinstance(flat, tuple) else flat, area_lut, background=(0.0, 0.0, 0.0, 0.0))


def test_off_origin_view_boxes_and_oversized_geometry():
    """View boxes that start off the origin and off the tile grid (negative, fractional), with geometry up to forty
    times larger than the view box: clipping to the view box, auxiliary fills and column backdrops above the rect."""
    rng = np.random.RandomState(5)
    for case in range(40):
        x0, y0 = float(rng.choice([0, 16, 37.5, -40, 1000])), float(rng.choice([0, 8, -33.25, 500]))
        w, h = float(rng.choice([33, 64, 100.5])), float(rng.choice([20, 64, 77.75]))
        b = SceneBuilderPy((x0, y0, x0 + w, y0 + h))
        for _ in range(int(rng.randint(1, 4))):
            s = 40.0 if rng.rand() < 0.2 else 1.0
            pt = lambda: (x0 + rng.uniform(-0.6, 1.6) * w * s, y0 + rng.uniform(-0.6, 1.6) * h * s)
            b.move_to(*pt())
            for _ in range(int(rng.randint(2, 6))):
                k = rng.randint(0, 3)
                if k == 0:
                    b.line_to(*pt())
                elif k == 1:
                    b.quad_to(*pt(), *pt())
                else:
                    b.cubic_to(*pt(), *pt(), *pt())
            b.close()
            b.end_path((int(rng.randint(0, 256)), 9, 9, 255 if rng.rand() < 0.5 else 100), int(rng.randint(0, 2)))
        compare(b.finish(f"adv{case}"))
