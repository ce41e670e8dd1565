"""Shared test plumbing: run the same FlatScene through the CPU oracle and the CUDA backend."""
from __future__ import annotations

import numpy as np

from oracle import pf_oracle as O
from pathfinder_b200.flat_scene import FlatScene


def oracle_scene(flat: FlatScene) -> O.OracleScene:
    # Paint ids as Scene::push_paint assigns them (equal colours share an id), like the CUDA-side Scene.
    paints, paint_colors = flat.palette()
    return O.make_scene(points=flat.points, point_flags=flat.point_flags, contour_offsets=flat.contour_offsets,
                        draw_contour_ranges=flat.contour_ranges(), draw_fill_rules=flat.fill_rules,
                        draw_paints=paints, paint_colors=paint_colors, view_box=flat.view_box,
                        clip_contour_ranges=flat.clip_contour_ranges if flat.n_clip_paths else None,
                        clip_fill_rules=flat.clip_fill_rules if flat.n_clip_paths else None,
                        draw_clip_paths=flat.draw_clip_paths if flat.n_clip_paths else None)


def oracle_options(xf=None, strip=None, dilation=(0.0, 0.0)):
    """xf = (m11, m12, m21, m22, tx, ty) as pathfinder_b200.scenes returns it."""
    t = None if xf is None else (xf[0], xf[2], xf[1], xf[3], xf[4], xf[5])
    return O.make_options(transform=t, strip=strip, dilation=dilation)


def oracle_build(flat: FlatScene, xf=None, strip=None, keep_lines=False, n_threads=1, dilation=(0.0, 0.0)) -> O.Built:
    return O.Built(oracle_scene(flat), oracle_options(xf, strip, dilation), n_threads=n_threads, keep_lines=keep_lines)


def cuda_render(flat: FlatScene, xf=None, size=None, background=None, debug=True, strip=None, renderer=None,
                dilation=(0.0, 0.0)):
    """Renders through the public API (Scene::build_and_render). Returns (renderer, image)."""
    from pathfinder_b200 import api
    w = int(size[0]) if size else int(flat.view_box[2])
    h = int(size[1]) if size else int(flat.view_box[3])
    r = renderer or api.CudaRenderer((w, h), background_color=background)
    r.set_debug_lists_enabled(debug)
    if strip is not None:
        r.set_strip(*strip)
    scene = api.Scene.from_flat(flat)
    t = None if xf is None else api.Transform2F(*xf)
    scene.build_and_render(r, api.BuildOptions(transform=t, dilation=dilation))
    return r, r.read_pixels()


def assert_records_equal(a: np.ndarray, b: np.ndarray, what: str):
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert len(a) == len(b), f"{what}: count {len(a)} != {len(b)}"
    if len(a) and a.tobytes() != b.tobytes():
        av, bv = a.view(np.uint8).reshape(len(a), -1), b.view(np.uint8).reshape(len(b), -1)
        bad = np.nonzero((av != bv).any(axis=1))[0]
        i = int(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {len(a)} records differ; first at {i}: {a[i]} != {b[i]}")


def live_tile_stats(built: O.Built):
    """Per framebuffer tile of the oracle's batch: number of list entries that survive the z-cull and the
    total number of fills behind them (what the fused fill+tile kernel has to walk there)."""
    tiles, z, rect = built.tiles, built.z_buffer, built.z_rect
    w = rect[2] - rect[0]
    inside = ((tiles["tile_x"] >= rect[0]) & (tiles["tile_x"] < rect[2]) &
              (tiles["tile_y"] >= rect[1]) & (tiles["tile_y"] < rect[3]))
    t = tiles[inside]
    fb = (t["tile_y"].astype(np.int64) - rect[1]) * w + (t["tile_x"].astype(np.int64) - rect[0])
    live = t["path_id"].astype(np.int64) >= z.reshape(-1)[fb]
    t, fb = t[live], fb[live]
    entries = np.bincount(fb, minlength=z.size)
    per_alpha = np.bincount(built.fills["link"], minlength=built.alpha_tile_count)
    alpha = t["alpha_tile_id"] != 0xFFFFFFFF
    per_tile = np.zeros(len(t), np.int64)
    per_tile[alpha] = per_alpha[t["alpha_tile_id"][alpha]]
    fills = np.bincount(fb, weights=per_tile, minlength=z.size).astype(np.int64)
    deepest_single = np.zeros(z.size, np.int64)
    np.maximum.at(deepest_single, fb, per_tile)
    return entries.reshape(z.shape), fills.reshape(z.shape), deepest_single.reshape(z.shape)


def interesting_crops(built: O.Built, frame_size, crop=512, n_random=5, seed=1):
    """Crop origins (pixels) for comparing a large frame with the oracle: around the framebuffer tile with the
    longest list, the one with the most fills, the one holding the alpha tile with the most fills, plus a few
    seeded random positions."""
    entries, fills, single = live_tile_stats(built)
    W, H = frame_size
    origins = []
    for grid in (entries, fills, single):
        ty, tx = np.unravel_index(int(grid.argmax()), grid.shape)
        origins.append((int(tx) * 16 + 8 - crop // 2, int(ty) * 16 + 8 - crop // 2))
    rng = np.random.default_rng(seed)
    for _ in range(n_random):
        origins.append((int(rng.integers(0, max(1, W - crop))), int(rng.integers(0, max(1, H - crop)))))
    clamp = lambda v, hi: max(0, min(v, hi))
    return [(clamp(x, max(0, W - crop)), clamp(y, max(0, H - crop))) for x, y in origins]


def assert_crops_match(built: O.Built, img: np.ndarray, area_lut, crops, crop=512, background=(0, 0, 0, 0), tol=1):
    """img (H x W x 4, RGBA8) against the oracle's composite on the given crops, within tol/255."""
    H_, W_ = img.shape[:2]
    for x0, y0 in crops:
        w, h = min(crop, W_ - x0), min(crop, H_ - y0)
        ref = built.render_crop(area_lut, (W_, H_), (x0, y0), (w, h), background=background)
        diff = np.abs(img[y0:y0 + h, x0:x0 + w].astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= tol, (f"crop at ({x0}, {y0}): max RGBA diff {diff.max()} at "
                                   f"{np.unravel_index(diff.argmax(), diff.shape)}")


def assert_lines_match(renderer, built: O.Built, n_paths: int):
    """Flattened lines, bit for bit and per path. The D3D11 builder drops paths whose tile rect is empty, the CPU
    tiler still flattens them: every group of GPU lines (one per kept path, in path order) must equal the oracle's
    lines of a distinct path, in order, and exactly path_count paths must be matched."""
    lines, paths = renderer.debug_lines()
    assert len(lines) == len(paths)
    starts = np.flatnonzero(np.r_[True, paths[1:] != paths[:-1]]) if len(paths) else np.zeros(0, np.int64)
    ends = np.r_[starts[1:], len(paths)]
    assert np.all(np.diff(paths[starts].astype(np.int64)) > 0), "lines are not grouped by ascending path"
    g = 0
    for p in range(n_paths):
        if g == len(starts):
            break
        ref = built.path_lines(p)
        mine = lines[starts[g]:ends[g]]
        if len(ref) == len(mine) and ref.view(np.uint32).tobytes() == mine.view(np.uint32).tobytes():
            g += 1
    assert g == len(starts), f"flattened lines differ: GPU path group {g} of {len(starts)} matches no oracle path"
    # (a kept path always has at least one line: closed contours, builder.rs:835)
