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


def compare(flat: FlatScene):
    built = H.oracle_build(flat, None, keep_lines=True)
    fills, tiles, lines = P.tile_scene(flat)
    for p in range(flat.n_paths):
        want = built.path_lines(p)
        got = np.asarray(lines[p], np.float32).reshape(-1, 4)
        assert want.tobytes() == got.tobytes(), f"path {p}: flattened lines differ"
    assert len(fills) == len(built.fills)
    got = np.asarray(fills, np.int64).reshape(-1, 5)
    for k, name in enumerate(("from_x", "from_y", "to_x", "to_y", "link")):
        assert np.array_equal(got[:, k], built.fills[name].astype(np.int64)), name
    # Every tile the oracle's batch lists (non-empty tiles that survive the z-buffer) carries the alpha-tile id and
    # the propagated backdrop the second tiler derives for that path and position.
    assert len(built.tiles) > 0 or len(fills) == 0
    for t in built.tiles:
        alpha, backdrop = tiles[int(t["path_id"])][(int(t["tile_x"]), int(t["tile_y"]))]
        assert alpha == int(t["alpha_tile_id"]) and backdrop == int(t["backdrop"]), t
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


def test_tiger_and_text_page():
    """The tiger at 128 px (transform applied to the points beforehand, in f32 with Transform2F's operation order)
    and a small text page: cubics and quadratics at realistic scales."""
    from tests.test_dilate_host import prepared_scene
    flat, xf = scenes.tiger(128)
    assert compare(prepared_scene(flat, xf, (0.0, 0.0))) > 1000
    assert compare(scenes.text_page(60, 128, layout="lines")) > 500
