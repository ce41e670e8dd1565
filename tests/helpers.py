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
