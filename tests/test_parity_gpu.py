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
        # (paths outside the view box are skipped by the D3D11 builder but still flattened by the CPU tiler)
        H.assert_lines_match(r, built, flat.n_paths)
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

    # The production configuration (no stage lists kept, fills of z-culled tiles never stored) must
    # give the same frame byte for byte, first in sizing mode and again from the cached batch.
    r2, img2 = H.cuda_render(flat, xf, size=size, background=background, debug=False)
    assert np.array_equal(img2, img), "production path differs from the instrumented path"
    scene_stats = r2.stats()
    assert scene_stats["fill_count"] == len(built.fills)
    r.close()
    r2.close()
    return img, built


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


def _star(b, cx, cy, r_out, r_in, points):
    """A closed star polygon with 2 * points short edges around (cx, cy)."""
    for i in range(2 * points):
        a = np.pi * i / points
        r = r_out if i % 2 == 0 else r_in
        x, y = cx + r * np.cos(a), cy + r * np.sin(a)
        b.move_to(x, y) if i == 0 else b.line_to(x, y)
    b.close()


@pytest.mark.parametrize("layers", [40, 200])
def test_deep_lists_and_dense_tiles(area_lut, layers):
    """The paths the headline scenes never reach (random100k@8192 tops out at 27 entries and 18 fills per
    tile): lists deeper than the in-shared-memory sort (32 entries) and than the round-1 cap (128), alpha tiles
    with more fills than one cooperative batch (32) and than the fast integer-to-float conversion covers (127),
    under both fill rules, and translucent solid layers interleaved with masks."""
    b = SceneBuilderPy((0, 0, 96, 64))
    rng = np.random.default_rng(layers)
    for i in range(layers):
        # translucent layers over the same few tiles: half of them whole-tile rectangles (solid entries), half
        # small shapes with edges (entries with fills), interleaved
        color = tuple(int(v) for v in rng.integers(0, 256, 3)) + (int(rng.integers(8, 64)),)
        if i % 2 == 0:
            x0, y0 = float(rng.uniform(-8, 20)), float(rng.uniform(-8, 12))
            b.move_to(x0, y0), b.line_to(x0 + 70, y0), b.line_to(x0 + 70, y0 + 50), b.line_to(x0, y0 + 50)
            b.close()
        else:
            cx, cy = float(rng.uniform(20, 70)), float(rng.uniform(16, 48))
            _star(b, cx, cy, float(rng.uniform(6, 20)), float(rng.uniform(2, 6)), int(rng.integers(3, 9)))
        b.end_path(color, FILL_RULE_EVEN_ODD if i % 3 == 0 else FILL_RULE_WINDING)
    # dense alpha tiles: stars with 40 and 180 short edges inside one 16-px tile
    _star(b, 40.0, 24.0, 7.5, 5.0, 20)
    b.end_path((10, 10, 200, 200), FILL_RULE_WINDING)
    _star(b, 56.0, 40.0, 7.8, 6.0, 90)
    b.end_path((200, 10, 10, 180), FILL_RULE_EVEN_ODD)
    flat = b.finish("deep")
    img, built = check_scene(flat, None, area_lut, background=(1.0, 1.0, 1.0, 1.0))
    entries, fills, single = H.live_tile_stats(built)
    assert entries.max() > (128 if layers >= 200 else 32), entries.max()
    assert single.max() > 127, single.max()


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


def test_cached_batch_and_strips_are_byte_identical(area_lut):
    """Steady state (batch cache hit, device-side counts) and the 4-strip partition reproduce the
    full single-pass frame exactly; the union of the strips' tile lists equals the full list."""
    from pathfinder_b200 import api
    flat = scenes.random_paths(1500, 512, 11, r_min=8.0, r_max=80.0)
    r = api.CudaRenderer((512, 512), background_color=(1, 1, 1, 1))
    scene = api.Scene.from_flat(flat)
    opts = api.BuildOptions()
    scene.build_and_render(r, opts)
    first = r.read_pixels()
    assert r.stats()["batch_cache_hits"] == 0 and r.stats()["host_sync_count"] >= 2
    scene.build_and_render(r, opts)
    second = r.read_pixels()
    s = r.stats()
    assert s["batch_cache_hits"] == 1 and s["host_sync_count"] == 1 and s["h2d_bytes"] == 0
    assert np.array_equal(first, second)
    # The steady state dices in one pass and appends lines in arbitrary order; the first frame of a
    # renderer (exact sizing) uses the ordered count -> scan -> emit path. Curves at a large scale
    # exercise deep subdivision: both paths must give the same pixels and the same totals.
    tflat, txf = scenes.tiger(1024)
    tr = api.CudaRenderer((1024, 1024), background_color=(1, 1, 1, 1))
    tscene = api.Scene.from_flat(tflat)
    topts = api.BuildOptions(transform=api.Transform2F(*txf))
    tscene.build_and_render(tr, topts)
    ordered, ordered_stats = tr.read_pixels(), tr.stats()
    for _ in range(2):
        tscene.build_and_render(tr, topts)
        streamed, streamed_stats = tr.read_pixels(), tr.stats()
        assert np.array_equal(ordered, streamed)
        for k in ("line_segment_count", "tile_list_entry_count", "visible_fill_count"):
            assert ordered_stats[k] == streamed_stats[k], k
    tr.close()

    full_tiles = None
    r.set_debug_lists_enabled(True)
    scene.build_and_render(r, opts)
    full_tiles = r.debug_tiles()
    stitched = np.zeros_like(first)
    parts = []
    for y0, y1 in [(0, 8), (8, 16), (16, 24), (24, 32)]:
        rs = api.CudaRenderer((512, 512), background_color=(1, 1, 1, 1))
        rs.set_debug_lists_enabled(True)
        rs.set_strip(y0, y1)
        scene.build_and_render(rs, opts)
        img = rs.read_pixels()
        stitched[y0 * 16:y1 * 16] = img[y0 * 16:y1 * 16]
        parts.append(rs.debug_tiles())
        rs.close()
    assert np.array_equal(stitched, first)
    key = lambda t: (int(t["path_id"]), int(t["tile_y"]), int(t["tile_x"]), int(t["backdrop"]), int(t["alpha_tile_id"] == 0xFFFFFFFF))
    assert sorted(key(t) for p in parts for t in p) == sorted(key(t) for t in full_tiles)
    r.close()


@pytest.mark.parametrize("deferred", [False, True])
def test_bound_overflow_is_repaired(area_lut, deferred):
    """A batch of the same shape as the previous one runs with the previous totals as buffer bounds. When
    the scene grew past them (same paths, larger transform) the frame must be re-rendered with exact
    sizes — at the end of the batch, or with deferred verification at the next use of the renderer — and
    the result must equal a fresh renderer's."""
    from pathfinder_b200 import api
    flat = scenes.random_paths(3000, 1024, 21, r_min=10.0, r_max=40.0)
    scene = api.Scene.from_flat(flat)
    small = api.BuildOptions(transform=api.Transform2F(0.1, 0.0, 0.0, 0.1, 100.0, 100.0))
    large = api.BuildOptions()
    r = api.CudaRenderer((1024, 1024), background_color=(1, 1, 1, 1))
    r.set_deferred_verification(deferred)
    scene.build_and_render(r, small)
    small_stats = r.stats()
    scene.build_and_render(r, large)
    img = r.read_pixels()
    s = r.stats()
    assert s["reruns"] == 1 and s["fill_count"] > 2 * small_stats["fill_count"]
    fresh = api.CudaRenderer((1024, 1024), background_color=(1, 1, 1, 1))
    scene.build_and_render(fresh, large)
    assert np.array_equal(img, fresh.read_pixels())
    assert fresh.stats()["fill_count"] == s["fill_count"] and fresh.stats()["reruns"] == 0
    # and the steady state afterwards needs neither a re-run nor, when deferred, a wait inside the frame
    scene.build_and_render(r, large)
    assert np.array_equal(r.read_pixels(), img) and r.stats()["reruns"] == 0
    fresh.close()
    r.close()


def test_wrong_level_commands_are_rejected():
    """Renderer::require_d3d11 (gpu/renderer.rs:1349-1360): D3D9 commands panic at the D3D11 level."""
    from pathfinder_b200 import _lib as L
    from pathfinder_b200 import api
    r = api.CudaRenderer((64, 64))
    r.begin_scene()
    for kind in (L.PF_RENDER_COMMAND_ADD_FILLS_D3D9, L.PF_RENDER_COMMAND_FLUSH_FILLS_D3D9, L.PF_RENDER_COMMAND_DRAW_TILES_D3D9):
        cmd = L.PFRenderCommand()
        cmd.kind = kind
        with pytest.raises(L.PathfinderCudaError) as e:
            r.render_command(cmd)
        assert e.value.status == L.PF_CUDA_ERROR_WRONG_LEVEL
    r.end_scene()
    with pytest.raises(L.PathfinderCudaError) as e:
        r.end_scene()
    assert e.value.status == L.PF_CUDA_ERROR_PROTOCOL
    with pytest.raises(L.PathfinderCudaError):
        api.CudaRenderer((64, 64), level=api.RendererLevel.D3D9)
    r.close()


# ---- BASELINE.json configurations at full size -------------------------------------------------

@pytest.mark.parametrize("even_odd", [False, True])
def test_full_size_tiger_4096(area_lut, even_odd):
    """configs[1]: tiger at 4096x4096, all-winding and odd-paths-even-odd. Lists bit-exact, RGBA
    within 1/255 of the oracle's evaluation of the reference fill math."""
    flat, xf = scenes.tiger(4096, even_odd_odd_paths=even_odd)
    built = H.oracle_build(flat, xf)
    r, img = H.cuda_render(flat, xf, background=(1.0, 1.0, 1.0, 1.0))
    H.assert_records_equal(r.debug_fills(), built.fills, "fills")
    H.assert_records_equal(r.debug_tiles(), built.tiles, "tiles")
    z, _ = r.debug_z_buffer()
    assert np.array_equal(z, built.z_buffer)
    ref = built.render(area_lut, 4096, 4096, background=(1.0, 1.0, 1.0, 1.0))
    assert np.abs(img.astype(np.int32) - ref.astype(np.int32)).max() <= RGBA_TOL
    r2, img2 = H.cuda_render(flat, xf, background=(1.0, 1.0, 1.0, 1.0), debug=False)
    assert np.array_equal(img, img2)
    r.close()
    r2.close()


def test_full_size_random100k_8192(area_lut):
    """configs[3]: 100k random cubic paths at 8192x8192. Fills, tiles and z-buffer bit-exact against
    the CPU tiler; RGBA within 1/255 of the oracle on eight 512x512 crops (around the longest list, the
    tile with the most fills, the densest alpha tile, and seeded random positions); the whole frame through
    size-independent properties (strips stitched = full frame, production path = instrumented path,
    deterministic across runs)."""
    from pathfinder_b200 import api
    flat = scenes.random_paths(100000, 8192, 0x5EED0004)
    built = H.oracle_build(flat, None)
    r, img = H.cuda_render(flat, None, background=(1.0, 1.0, 1.0, 1.0))
    s = r.stats()
    assert s["line_segment_count"] == built.line_segment_count
    assert s["input_segment_count"] == built.input_segment_count
    H.assert_records_equal(r.debug_fills(), built.fills, "fills")
    H.assert_records_equal(r.debug_tiles(), built.tiles, "tiles")
    z, _ = r.debug_z_buffer()
    assert np.array_equal(z, built.z_buffer)
    assert s["alpha_tile_count"] == built.alpha_tile_count
    r.close()
    H.assert_crops_match(built, img, area_lut, H.interesting_crops(built, (8192, 8192)),
                         background=(1.0, 1.0, 1.0, 1.0), tol=RGBA_TOL)
    # production path, twice (sizing frame, then cached batch), and two strips
    rp = api.CudaRenderer((8192, 8192), background_color=(1.0, 1.0, 1.0, 1.0))
    scene = api.Scene.from_flat(flat)
    opts = api.BuildOptions()
    scene.build_and_render(rp, opts)
    a = rp.read_pixels()
    scene.build_and_render(rp, opts)
    b = rp.read_pixels()
    assert np.array_equal(a, img) and np.array_equal(b, img)
    for y0, y1 in [(0, 256), (256, 512)]:
        rp.set_strip(y0, y1)
        scene.build_and_render(rp, opts)
        part = rp.read_pixels()
        assert np.array_equal(part[y0 * 16:y1 * 16], img[y0 * 16:y1 * 16])
    rp.close()


def test_full_size_random1m_16384(area_lut):
    """configs[4]: 1M random cubic paths at 16384x16384 (the multi-GPU configuration), on one GPU. Tiles, z-buffer
    and the fill / alpha-tile / line totals bit-exact against the CPU tiler (the 120M-record fill list itself is
    compared at 100k paths above); RGBA within 1/255 of the oracle on eight 512x512 crops (deepest lists
    included); the 1 GiB frame through size-independent properties: production path = instrumented path,
    cached batch = first frame, two of the eight strips = the same rows of the full frame."""
    from pathfinder_b200 import api
    size = 16384
    flat = scenes.random_paths(1000000, size, 0x5EED0005)
    built = H.oracle_build(flat, None)
    r, img = H.cuda_render(flat, None, background=(1.0, 1.0, 1.0, 1.0))
    s = r.stats()
    assert s["line_segment_count"] == built.line_segment_count
    assert s["input_segment_count"] == built.input_segment_count
    assert s["fill_count"] == len(built.fills) and s["alpha_tile_count"] == built.alpha_tile_count
    H.assert_records_equal(r.debug_tiles(), built.tiles, "tiles")
    z, _ = r.debug_z_buffer()
    assert np.array_equal(z, built.z_buffer)
    r.close()
    H.assert_crops_match(built, img, area_lut, H.interesting_crops(built, (size, size)),
                         background=(1.0, 1.0, 1.0, 1.0), tol=RGBA_TOL)
    del built
    rp = api.CudaRenderer((size, size), background_color=(1.0, 1.0, 1.0, 1.0))
    scene = api.Scene.from_flat(flat)
    opts = api.BuildOptions()
    for _ in range(2):
        scene.build_and_render(rp, opts)
        assert np.array_equal(rp.read_pixels(), img)
    rows = size // 16 // 8
    for g in (0, 5):
        rp.set_strip(g * rows, (g + 1) * rows)
        scene.build_and_render(rp, opts)
        part = rp.read_pixels()
        assert np.array_equal(part[g * rows * 16:(g + 1) * rows * 16], img[g * rows * 16:(g + 1) * rows * 16])
    rp.close()


def test_edge_inputs(area_lut):
    """Empty and ragged inputs, degenerate geometry, extreme coordinates, maximum winding depth."""
    b = SceneBuilderPy((0, 0, 160, 96))
    b.end_path((9, 9, 9, 255))                      # empty path
    b.move_to(30, 30)
    b.close()
    b.end_path((9, 9, 9, 255))                      # one point
    b.move_to(5, 5)
    b.line_to(150, 5)
    b.close()
    b.end_path((9, 9, 9, 255))                      # zero area: fills that cancel
    b.move_to(-1.0e6, -2.0e6)
    b.line_to(3.0e6, 40)
    b.line_to(80, 4.0e6)
    b.close()
    b.end_path((30, 60, 200, 180))                  # huge triangle clipped on every side
    b.move_to(16, 16)
    b.line_to(48, 16)
    b.line_to(48, 48)
    b.line_to(16, 48)
    b.close()
    b.end_path((200, 10, 10, 255))                  # edges exactly on tile boundaries
    for _ in range(140):                            # > 127 nested windings: i8 backdrop wraps (quirk 7)
        b.move_to(100, 20)
        b.line_to(150, 20)
        b.line_to(150, 80)
        b.line_to(100, 80)
        b.close()
    b.end_path((10, 200, 10, 200))
    b.move_to(60, 60)
    b.cubic_to(60, 60, 60, 60, 60, 60)              # degenerate cubic
    b.quad_to(90, 90, 60, 90)
    b.close()
    b.end_path((0, 0, 0, 128), FILL_RULE_EVEN_ODD)
    check_scene(b.finish("edges"), None, area_lut)


# ---- clip paths -----------------------------------------------------------------

def clip_scene(size=256, n_draw=40, seed=5):
    """Draw paths clipped by two clip paths (a curved blob with a hole, even-odd; a triangle, winding), plus
    unclipped ones, opaque and translucent: exercises all four cases of Tiler::prepare_tiles — both masks,
    solid draw tile under a clip mask, tile outside the clip path, tile fully inside it."""
    rng = np.random.RandomState(seed)
    b = SceneBuilderPy((0, 0, size, size))
    s = size / 256.0
    b.move_to(128 * s, 20 * s)
    b.quad_to(236 * s, 20 * s, 236 * s, 128 * s)
    b.cubic_to(236 * s, 200 * s, 190 * s, 236 * s, 128 * s, 236 * s)
    b.quad_to(20 * s, 236 * s, 20 * s, 128 * s)
    b.quad_to(20 * s, 20 * s, 128 * s, 20 * s)
    b.close()
    b.move_to(100 * s, 100 * s)
    b.line_to(160 * s, 100 * s)
    b.line_to(160 * s, 160 * s)
    b.line_to(100 * s, 160 * s)
    b.close()
    blob = b.end_clip_path(FILL_RULE_EVEN_ODD)
    b.move_to(10 * s, 240 * s)
    b.line_to(250 * s, 200 * s)
    b.line_to(90 * s, 5 * s)
    b.close()
    tri = b.end_clip_path(FILL_RULE_WINDING)
    # a full-frame opaque rectangle under the blob: solid draw tiles x clip masks (the "replace" case)
    b.move_to(0, 0)
    b.line_to(size, 0)
    b.line_to(size, size)
    b.line_to(0, size)
    b.close()
    b.end_path((30, 60, 200, 255), clip=blob)
    for i in range(n_draw):
        cx, cy = rng.uniform(0, size, 2)
        r = rng.uniform(10, 70) * s
        k = rng.randint(3, 7)
        for m in range(k):
            a = 2 * np.pi * m / k + rng.uniform(-0.2, 0.2)
            x, y = cx + r * np.cos(a) * rng.uniform(0.6, 1.0), cy + r * np.sin(a) * rng.uniform(0.6, 1.0)
            if m == 0:
                b.move_to(x, y)
            elif m % 2:
                b.quad_to(cx + 1.3 * r * np.cos(a - 0.5), cy + 1.3 * r * np.sin(a - 0.5), x, y)
            else:
                b.line_to(x, y)
        b.close()
        colour = tuple(int(v) for v in rng.randint(0, 256, 3)) + ((255,) if i % 2 else (int(rng.randint(60, 250)),))
        b.end_path(colour, FILL_RULE_EVEN_ODD if i % 3 == 0 else FILL_RULE_WINDING,
                   clip=(blob, tri, 0xFFFFFFFF)[i % 3])
    return b.finish("clips")


@pytest.mark.parametrize("size", [256, 1024])
def test_clip_paths(area_lut, size):
    """A clipped scene against the oracle's CPU tiler + D3D9 clip combine (tiler.rs:114-156,
    tile_clip_combine.fs.glsl:28-31): fills (clip paths first), tiles, Clip records and z-buffer bit-exact,
    pixels within 1/255. Both clip paths are used, in scene order, so the D3D11 clip batch (used clip paths in
    order of first use) numbers alpha tiles like the CPU tiler (all clip paths in scene order)."""
    from pathfinder_b200 import api
    flat = clip_scene(size)
    built = H.oracle_build(flat, None)
    assert len(built.clips) > 0
    ref = built.render(area_lut, size, size, background=(1.0, 1.0, 1.0, 1.0))
    rd, img_debug = H.cuda_render(flat, None, size=(size, size), background=(1.0, 1.0, 1.0, 1.0), debug=True)
    H.assert_records_equal(rd.debug_fills(), built.fills, "fills")
    H.assert_records_equal(rd.debug_tiles(), built.tiles, "tiles")
    H.assert_records_equal(rd.debug_clips(), built.clips, "clips")
    z, rect = rd.debug_z_buffer()
    assert rect == built.z_rect and np.array_equal(z, built.z_buffer)
    st = rd.stats()
    assert st["alpha_tile_count"] == built.alpha_tile_count and st["fill_count"] == len(built.fills)
    r, img = H.cuda_render(flat, None, size=(size, size), background=(1.0, 1.0, 1.0, 1.0), debug=False)
    diff = np.abs(img.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= RGBA_TOL, f"max RGBA diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}"
    assert np.array_equal(img_debug, img), "production path differs from the instrumented path"
    rd.close()
    # the clip really matters: the same draw paths without their clip paths render differently
    import dataclasses
    unclipped = dataclasses.replace(flat, draw_clip_paths=np.full(flat.n_paths, 0xFFFFFFFF, np.uint32))
    r0, img0 = H.cuda_render(unclipped, None, size=(size, size), background=(1.0, 1.0, 1.0, 1.0), debug=False)
    assert np.abs(img0.astype(np.int32) - img.astype(np.int32)).max() > 100
    # a second frame of the same renderer, and 2 strips, reproduce the frame exactly
    scene = api.Scene.from_flat(flat)
    scene.build_and_render(r, api.BuildOptions())
    assert np.array_equal(r.read_pixels(), img)
    rows = size // 16
    stitched = np.zeros_like(img)
    for y0, y1 in [(0, rows // 2), (rows // 2, rows)]:
        rs, part = H.cuda_render(flat, None, size=(size, size), background=(1.0, 1.0, 1.0, 1.0), debug=False, strip=(y0, y1))
        stitched[y0 * 16:y1 * 16] = part[y0 * 16:y1 * 16]
        rs.close()
    assert np.array_equal(stitched, img)
    r.close()
    r0.close()


def test_deferred_verification_and_accumulated_times(area_lut):
    """Two renderers sharing one stream with deferred verification: frames are enqueued without a host wait,
    the accumulated stage times cover every batch, and the pixels equal the synchronous frames."""
    import torch
    from pathfinder_b200 import api
    stream = torch.cuda.Stream()
    items = []
    for flat, xf, size in [(scenes.tiger(512)[0], scenes.tiger(512)[1], 512), (scenes.random_paths(2000, 1024, 3), None, 1024)]:
        r = api.CudaRenderer((size, size), background_color=(1, 1, 1, 1))
        r.set_stream(stream.cuda_stream)
        scene = api.Scene.from_flat(flat)
        opts = api.BuildOptions(transform=None if xf is None else api.Transform2F(*xf))
        scene.build_and_render(r, opts)
        items.append((r, scene, opts, r.read_pixels()))
    for r, *_ in items:
        r.set_deferred_verification(True)
        r.set_timing_enabled(True)
    for _ in range(5):
        for r, scene, opts, _ in items:
            scene.build_and_render(r, opts)
    for r, scene, opts, ref in items:
        totals, batches = r.accumulated_times()
        assert batches == 5 and totals["total_ms"] > 0 and totals["fill_tile_ms"] > 0
        assert abs(sum(totals[k] for k in ("bound_ms", "dice_ms", "bin_ms", "propagate_ms", "sort_ms", "fill_tile_ms")) - totals["total_ms"]) < 0.05 * totals["total_ms"]
        s = r.stats()
        assert s["reruns"] == 0 and s["host_sync_count"] <= 1
        assert np.array_equal(r.read_pixels(), ref)
        r.close()


# ---- seeded fuzz: many small scenes of mixed geometry against the oracle --------------------------

def fuzz_scene(seed, rotate=False):
    """A small random scene: lines, quadratics and cubics mixed inside each contour, several contours per path,
    coordinates that stray outside the view box, points exactly on tile and pixel boundaries, tiny and huge
    shapes, opaque and translucent paints, both fill rules; view boxes that are not a multiple of the tile size."""
    rng = np.random.RandomState(seed)
    w, h = int(rng.choice([37, 64, 100, 160, 250])), int(rng.choice([33, 64, 96, 130, 200]))
    ox, oy = 0.0, 0.0  # the destination image starts at the origin, like the reference's viewport
    b = SceneBuilderPy((ox, oy, ox + w, oy + h))
    snap = lambda v: float(np.round(v * rng.choice([1.0, 1.0, 1.0 / 16.0, 4.0])) / rng.choice([1.0, 16.0, 4.0])) if rng.rand() < 0.15 else float(v)

    def point(cx, cy, r):
        return snap(cx + rng.uniform(-r, r)), snap(cy + rng.uniform(-r, r))

    for _ in range(int(rng.randint(1, 14))):
        cx, cy = ox + rng.uniform(-0.3, 1.3) * w, oy + rng.uniform(-0.3, 1.3) * h
        r = float(np.exp(rng.uniform(np.log(0.3), np.log(1.5 * max(w, h)))))
        for _ in range(int(rng.randint(1, 4))):
            b.move_to(*point(cx, cy, r))
            for _ in range(int(rng.randint(1, 7))):
                kind = rng.randint(0, 3)
                if kind == 0:
                    b.line_to(*point(cx, cy, r))
                elif kind == 1:
                    b.quad_to(*point(cx, cy, 2 * r), *point(cx, cy, r))
                else:
                    b.cubic_to(*point(cx, cy, 2 * r), *point(cx, cy, 2 * r), *point(cx, cy, r))
            b.close()
        alpha = 255 if rng.rand() < 0.5 else int(rng.randint(1, 255))
        b.end_path(tuple(int(v) for v in rng.randint(0, 256, 3)) + (alpha,),
                   FILL_RULE_EVEN_ODD if rng.rand() < 0.5 else FILL_RULE_WINDING)
    xf = None
    if seed % 2:
        s = float(np.exp(rng.uniform(np.log(0.25), np.log(6.0))))
        xf = (s, 0.0, 0.0, s, float(rng.uniform(-0.5, 0.5) * w), float(rng.uniform(-0.5, 0.5) * h))
    if rotate:
        # a general affine map (rotation, anisotropic scale, shear) about the middle of the view box
        a = float(rng.uniform(0, 2 * np.pi))
        sx, sy, k = float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.5, 2.0)), float(rng.uniform(-0.5, 0.5))
        m11, m12 = sx * np.cos(a), -sy * np.sin(a) + k * sx * np.cos(a)
        m21, m22 = sx * np.sin(a), sy * np.cos(a) + k * sx * np.sin(a)
        cx, cy = 0.5 * w, 0.5 * h
        xf = tuple(float(np.float32(v)) for v in (m11, m12, m21, m22, cx - (m11 * cx + m12 * cy), cy - (m21 * cx + m22 * cy)))
    return b.finish(f"fuzz{seed}"), xf, (w, h)


@pytest.mark.parametrize("block", range(4))
def test_fuzz_small_scenes(area_lut, block):
    """25 seeded random scenes per block: lines, fills (with alpha tile ids), tiles, z-buffer bit-exact, masks and
    pixels within 1/255, production path identical to the instrumented one (check_scene)."""
    for seed in range(block * 25, block * 25 + 25):
        flat, xf, (w, h) = fuzz_scene(seed)
        try:
            background = (0.9, 0.95, 1.0, 1.0) if seed % 4 else None
            img, _ = check_scene(flat, xf, area_lut, size=(w, h), background=background)
            if seed % 5 == 0:  # three uneven strips of tile rows reproduce the frame
                rows = (h + 15) // 16
                cuts = sorted({0, rows // 3, (2 * rows + 2) // 3, rows})
                stitched = np.zeros_like(img)
                for y0, y1 in zip(cuts[:-1], cuts[1:]):
                    rs, part = H.cuda_render(flat, xf, size=(w, h), background=background, debug=False, strip=(y0, y1))
                    stitched[y0 * 16:y1 * 16] = part[y0 * 16:y1 * 16]
                    rs.close()
                assert np.array_equal(stitched, img), "strips differ from the full frame"
        except AssertionError as e:
            raise AssertionError(f"fuzz seed {seed}: {e}") from e


def test_fuzz_general_affine_transforms(area_lut):
    """BuildOptions transforms with rotation and shear: the D3D11 builder's tile rects (transformed bounding box)
    are then a superset of the CPU tiler's (bounds of the transformed outline); the extra tiles are empty, so
    fills, non-empty tiles, z-buffer and pixels still match. Flattened lines are compared too: both sides
    transform the control points before dicing."""
    for seed in range(200, 230):
        flat, xf, (w, h) = fuzz_scene(seed, rotate=True)
        try:
            check_scene(flat, xf, area_lut, size=(w, h), background=(1.0, 1.0, 1.0, 1.0))
        except AssertionError as e:
            raise AssertionError(f"affine fuzz seed {seed}: {e}") from e


def fuzz_clip_scene(seed):
    """Like fuzz_scene, with one or two clip paths (used in scene order, see test_clip_paths) and draw paths that
    are clipped by them or not."""
    rng = np.random.RandomState(1000 + seed)
    w, h = int(rng.choice([64, 100, 160, 250])), int(rng.choice([64, 96, 130, 200]))
    b = SceneBuilderPy((0.0, 0.0, w, h))

    def blob(cx, cy, r, curvy):
        k = int(rng.randint(3, 8))
        for m in range(k):
            a = 2 * np.pi * m / k + rng.uniform(-0.3, 0.3)
            x, y = cx + r * np.cos(a) * rng.uniform(0.5, 1.0), cy + r * np.sin(a) * rng.uniform(0.5, 1.0)
            if m == 0:
                b.move_to(x, y)
            elif curvy and m % 2:
                b.quad_to(cx + 1.4 * r * np.cos(a - 0.4), cy + 1.4 * r * np.sin(a - 0.4), x, y)
            else:
                b.line_to(x, y)
        b.close()

    clips = []
    for _ in range(int(rng.randint(1, 3))):
        for _ in range(int(rng.randint(1, 3))):  # one or two contours (holes / islands with even-odd)
            blob(rng.uniform(0.2, 0.8) * w, rng.uniform(0.2, 0.8) * h, rng.uniform(0.15, 0.6) * max(w, h), True)
        clips.append(b.end_clip_path(FILL_RULE_EVEN_ODD if rng.rand() < 0.5 else FILL_RULE_WINDING))
    n_draw = int(rng.randint(len(clips), 12))
    for i in range(n_draw):
        big = rng.rand() < 0.3
        r = (rng.uniform(0.5, 1.2) if big else rng.uniform(0.05, 0.4)) * max(w, h)
        blob(rng.uniform(-0.1, 1.1) * w, rng.uniform(-0.1, 1.1) * h, r, rng.rand() < 0.6)
        alpha = 255 if rng.rand() < 0.5 else int(rng.randint(1, 255))
        clip = clips[i] if i < len(clips) else (int(rng.choice(clips)) if rng.rand() < 0.6 else 0xFFFFFFFF)
        b.end_path(tuple(int(v) for v in rng.randint(0, 256, 3)) + (alpha,),
                   FILL_RULE_EVEN_ODD if rng.rand() < 0.4 else FILL_RULE_WINDING, clip=clip)
    return b.finish(f"clipfuzz{seed}"), (w, h)


def test_fuzz_clipped_scenes(area_lut):
    """40 seeded random clipped scenes: fills (clip paths first), tiles, Clip records and z-buffer bit-exact,
    pixels within 1/255, production frame identical to the instrumented one."""
    for seed in range(40):
        flat, (w, h) = fuzz_clip_scene(seed)
        built = H.oracle_build(flat, None)
        try:
            rd, img = H.cuda_render(flat, None, size=(w, h), background=(1.0, 1.0, 1.0, 1.0), debug=True)
            H.assert_records_equal(rd.debug_fills(), built.fills, "fills")
            H.assert_records_equal(rd.debug_tiles(), built.tiles, "tiles")
            H.assert_records_equal(rd.debug_clips(), built.clips, "clips")
            z, rect = rd.debug_z_buffer()
            assert rect == built.z_rect and np.array_equal(z, built.z_buffer), "z-buffer"
            ref = built.render(area_lut, w, h, background=(1.0, 1.0, 1.0, 1.0))
            diff = np.abs(img.astype(np.int32) - ref.astype(np.int32))
            assert diff.max() <= RGBA_TOL, f"max RGBA diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}"
            rp, img2 = H.cuda_render(flat, None, size=(w, h), background=(1.0, 1.0, 1.0, 1.0), debug=False)
            assert np.array_equal(img2, img), "production path differs from the instrumented path"
            rd.close()
            rp.close()
        except AssertionError as e:
            raise AssertionError(f"clip fuzz seed {seed}: {e}") from e
