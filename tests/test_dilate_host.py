"""Outline dilation on the host (stem darkening): dilate.cpp behind PFOutlineDilate and its use by
PFSceneBuild when the build options carry a dilation. Checked through geometric properties and, bit for bit, against
the oracle's restatement of Scene::apply_render_options (transform, then dilate) — the two were written separately
(oracle/pf_oracle.cpp for the checker, csrc/dilate.cpp + csrc/scene.cpp for the product)."""
import ctypes as C

import numpy as np
import pytest

from pathfinder_b200 import _lib as L
from pathfinder_b200 import api, scenes
from pathfinder_b200.flat_scene import FlatScene, SceneBuilderPy
from tests import helpers as H
from tests.test_scene_host import collect

R = np.float32(1.0) / np.sqrt(np.float32(2.0))


def test_square_grows_along_its_diagonals_either_winding():
    """Each corner moves along the bisector of its edges by `amount` per axis times the bisector's unit components:
    a square grows by amount/sqrt(2) per side, whichever way it is wound (Orientation::from_outline decides)."""
    square = np.array([(10, 10), (20, 10), (20, 20), (10, 20)], np.float32)
    want = np.array([(10 - R, 10 - R), (20 + R, 10 - R), (20 + R, 20 + R), (10 - R, 20 + R)], np.float32)
    got = api.dilate_outline(square, [0, 4], (1.0, 1.0))
    assert np.allclose(got, want, atol=1e-6)
    got = api.dilate_outline(square[::-1], [0, 4], (1.0, 1.0))
    assert np.allclose(got, want[::-1], atol=1e-6)
    # anisotropic amounts scale the axes separately
    got = api.dilate_outline(square, [0, 4], (2.0, 0.5))
    assert np.allclose(got[0], (10 - 2 * R, 10 - 0.5 * R), atol=1e-6)


def test_hole_shrinks_when_the_outer_contour_grows():
    """A glyph-like outline: the outer contour decides the orientation, so the oppositely wound hole closes up."""
    outer = [(0, 0), (30, 0), (30, 30), (0, 30)]
    hole = [(10, 10), (10, 20), (20, 20), (20, 10)]
    got = api.dilate_outline(np.array(outer + hole, np.float32), [0, 4, 8], (1.0, 1.0))
    assert np.allclose(got[0], (-R, -R), atol=1e-6) and np.allclose(got[2], (30 + R, 30 + R), atol=1e-6)
    assert np.allclose(got[4], (10 + R, 10 + R), atol=1e-6) and np.allclose(got[6], (20 - R, 20 - R), atol=1e-6)


def test_coincident_points_move_together_and_degenerate_contours_stay():
    pts = np.array([(10, 10), (10, 10), (20, 10), (20, 20), (20, 20), (10, 20), (10, 10)], np.float32)
    got = api.dilate_outline(pts, [0, 7], (1.0, 1.0))
    assert np.array_equal(got[0], got[1]) and np.array_equal(got[0], got[6]) and np.array_equal(got[3], got[4])
    assert np.allclose(got[0], (10 - R, 10 - R), atol=1e-6) and np.allclose(got[3], (20 + R, 20 + R), atol=1e-6)
    # a contour of one position has no edges: it is left alone (the run never closes, dilation.rs:107-111)
    dot = np.array([(5, 5), (5, 5), (5, 5)], np.float32)
    assert np.array_equal(api.dilate_outline(dot, [0, 3], (1.0, 1.0)), dot)
    # zero amount is a no-op (Vector2F::is_zero, scene.rs:268), as are empty inputs
    assert np.array_equal(api.dilate_outline(pts, [0, 7], (0.0, 0.0)), pts)
    assert len(api.dilate_outline(np.zeros((0, 2), np.float32), [0], (1.0, 1.0))) == 0


def test_straight_through_point_moves_along_the_normal():
    """A point in the middle of a straight edge has bisector = the edge normal: it moves by the full amount."""
    pts = np.array([(0, 0), (10, 0), (20, 0), (20, 10), (0, 10)], np.float32)
    got = api.dilate_outline(pts, [0, 5], (1.0, 1.0))
    assert np.allclose(got[1], (10, -1), atol=1e-6)


def prepared_scene(flat: FlatScene, xf, dilation) -> FlatScene:
    """What the CPU tiler tiles for (flat, xf, dilation): points transformed in f32 with Transform2F's operation order
    ((m11 x + m12 y) + tx), then every path's outline dilated through the C ABI."""
    f = np.float32
    pts = np.asarray(flat.points, np.float32)
    if xf is not None:
        m11, m12, m21, m22, tx, ty = [f(v) for v in xf]
        x, y = pts[:, 0].copy(), pts[:, 1].copy()
        pts = np.stack([(m11 * x + m12 * y) + tx, (m21 * x + m22 * y) + ty], axis=1).astype(np.float32)
    out = pts.copy()
    co = np.asarray(flat.contour_offsets, np.int64)
    ranges = np.concatenate([flat.contour_ranges(), flat.clip_contour_ranges]).astype(np.int64)
    for c0, c1 in ranges:  # draw paths, then clip paths
        if c0 == c1:
            continue
        p0, p1 = co[c0], co[c1]
        out[p0:p1] = api.dilate_outline(pts[p0:p1], co[c0:c1 + 1] - p0, dilation)
    clips = flat.n_clip_paths > 0
    return FlatScene(out, flat.point_flags, flat.contour_offsets, flat.path_contour_offsets, flat.fill_rules, flat.paints,
                     flat.paint_colors, flat.view_box, flat.name + "/prepared",
                     clip_contour_ranges=flat.clip_contour_ranges if clips else None,
                     clip_fill_rules=flat.clip_fill_rules if clips else None,
                     draw_clip_paths=flat.draw_clip_paths if clips else None)


def clipped_scene() -> FlatScene:
    b = SceneBuilderPy((0, 0, 128, 128))
    b.move_to(32, 32); b.line_to(96, 32); b.line_to(96, 96); b.line_to(32, 96); b.close()
    clip = b.end_clip_path()
    b.move_to(10, 64); b.quad_to(64, -20, 118, 64); b.quad_to(64, 150, 10, 64); b.close()
    b.end_path((200, 30, 30, 255), clip=clip)
    b.move_to(5, 5); b.line_to(60, 8); b.line_to(30, 70); b.close()
    b.end_path((30, 30, 200, 160))
    return b.finish("clipped")


def scene_cases():
    tiger, xf = scenes.tiger(256)
    yield "tiger@256", tiger, xf, (0.35, 0.2)
    yield "text", scenes.text_page(600, 512, layout="lines"), None, (0.3, 0.3)
    yield "clipped", clipped_scene(), (0.9, 0.1, -0.1, 0.9, 6.0, 4.0), (1.0, 0.75)  # rotation + shear, clip path
    # anisotropic scale + translation over quadratic outlines, stem-darkening-sized amounts (effects.rs:32)
    yield "text/scaled", scenes.text_page(200, 512), (1.5, 0.0, 0.0, 0.75, 8.0, -3.0), (0.0121 * 16, 0.0121 * 1.25 * 16)


@pytest.mark.parametrize("name,flat,xf,dilation", list(scene_cases()), ids=lambda v: v if isinstance(v, str) else None)
def test_host_dilation_matches_the_cpu_tiler(name, flat, xf, dilation):
    """Oracle(original scene, transform + dilation options) == Oracle(points prepared by the product's host code, no
    options): flattened lines, fills and tile lists bit for bit."""
    want = H.oracle_build(flat, xf, keep_lines=True, dilation=dilation)
    got = H.oracle_build(prepared_scene(flat, xf, dilation), None, keep_lines=True)
    assert want.line_segment_count == got.line_segment_count and want.line_segment_count > 0
    for p in range(flat.n_clip_paths + flat.n_paths):
        a, b = want.path_lines(p), got.path_lines(p)
        assert a.tobytes() == b.tobytes(), f"{name}: path {p} flattens differently"
    H.assert_records_equal(want.fills, got.fills, f"{name} fills")
    plain = H.oracle_build(flat, xf)
    assert want.fills.tobytes() != plain.fills.tobytes()  # the dilation did something


def uploaded_points(flat: FlatScene, prepared: FlatScene) -> np.ndarray:
    """SegmentsD3D11::add_path layout of a scene's points: every contour followed by its first point again."""
    co = np.asarray(flat.contour_offsets, np.int64)
    out = []
    for c in range(len(co) - 1):
        if co[c] == co[c + 1]:
            continue
        out.append(prepared.points[co[c]:co[c + 1]])
        out.append(prepared.points[co[c]:co[c] + 1])
    return np.concatenate(out).astype(np.float32)


def test_scene_build_applies_dilation_after_the_transform():
    flat, xf = scenes.tiger(256)
    dilation = (0.35, 0.2)
    scene = api.Scene.from_flat(flat)
    sink = L.PFSceneSinkState()
    cmds = collect(scene, api.BuildOptions(transform=api.Transform2F(*xf), dilation=dilation), sink)
    kinds = [c["kind"] for c in cmds]
    assert kinds == ["Start", "UploadTextureMetadata", "UploadSceneD3D11", "DrawTilesD3D11", "Finish"]
    upload, draw = cmds[2], cmds[3]
    want = uploaded_points(flat, prepared_scene(flat, xf, dilation))
    assert upload["points"].tobytes() == want.tobytes()
    # the device dices prepared points under an identity transform, inside the CPU tiler's tile rects
    assert draw["transform"] == (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
    built = H.oracle_build(flat, xf, dilation=dilation)
    assert draw["tile_count"] == built.bbox_tile_count
    assert draw["segment_count"] == built.input_segment_count

    # same options again: nothing to upload, same batch key
    again = collect(scene, api.BuildOptions(transform=api.Transform2F(*xf), dilation=dilation), sink)
    assert [c["kind"] for c in again] == ["Start", "UploadTextureMetadata", "DrawTilesD3D11", "Finish"]
    assert again[2]["content_key"] == draw["content_key"]
    # another amount: the prepared points change, so the scene is uploaded again under a new key
    other = collect(scene, api.BuildOptions(transform=api.Transform2F(*xf), dilation=(0.5, 0.5)), sink)
    assert "UploadSceneD3D11" in [c["kind"] for c in other]
    assert other[3]["content_key"] != draw["content_key"]
    assert other[2]["points"].tobytes() != want.tobytes()
    # no dilation: back to the scene's own points and the transform on the device
    plain = collect(scene, api.BuildOptions(transform=api.Transform2F(*xf)), sink)
    assert "UploadSceneD3D11" in [c["kind"] for c in plain]
    assert plain[2]["points"].tobytes() == uploaded_points(flat, flat).tobytes()
    assert plain[3]["transform"] == tuple(np.float32(v) for v in xf)
    assert plain[3]["tile_count"] == H.oracle_build(flat, xf).bbox_tile_count


def test_two_sinks_each_get_the_prepared_scene():
    """A second renderer (its own SceneSink) building the same scene with the same options still gets its upload."""
    flat = scenes.text_page(100, 256)
    scene = api.Scene.from_flat(flat)
    options = api.BuildOptions(dilation=(0.2, 0.2))
    a, b = L.PFSceneSinkState(), L.PFSceneSinkState()
    first = collect(scene, options, a)
    second = collect(scene, options, b)
    assert "UploadSceneD3D11" in [c["kind"] for c in first] and "UploadSceneD3D11" in [c["kind"] for c in second]
    assert first[2]["points"].tobytes() == second[2]["points"].tobytes()
    assert "UploadSceneD3D11" not in [c["kind"] for c in collect(scene, options, a)]


def test_clip_paths_are_dilated_too():
    """build_clip_path_on_cpu runs the clip outline through apply_render_options as well (builder.rs:226-247): the
    clip segments are the prepared ones and the clip batch's tile rect is the dilated bounds."""
    flat = clipped_scene()
    seen = {}

    def listener(cmd):
        if cmd.kind == L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11:
            cs = cmd.u.upload_scene_d3d11.clip_segments
            seen["clip_points"] = np.ctypeslib.as_array(C.cast(cs.points, C.POINTER(C.c_float)), (cs.point_count, 2)).copy()
        if cmd.kind == L.PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11:
            batch = cmd.u.prepare_clip_tiles_d3d11.batch
            pm = C.cast(batch.prepare_info.propagate_metadata, C.POINTER(L.PFPropagateMetadataD3D11))[0]
            seen["rect"] = (pm.tile_rect.origin.x, pm.tile_rect.origin.y, pm.tile_rect.lower_right.x, pm.tile_rect.lower_right.y)
            t = batch.prepare_info.transform
            seen["transform"] = (t.matrix.m00, t.matrix.m01, t.matrix.m10, t.matrix.m11, t.vector.x, t.vector.y)

    api.Scene.from_flat(flat).build(api.BuildOptions(dilation=(1.0, 1.0)), listener)
    assert seen["rect"] == (1, 1, 7, 7)  # bounds (32..96) grown by the amount = 31..97 px
    assert seen["transform"] == (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
    want = np.array([(32 - R, 32 - R), (96 + R, 32 - R), (96 + R, 96 + R), (32 - R, 96 + R), (32 - R, 32 - R)], np.float32)
    assert np.allclose(seen["clip_points"], want, atol=1e-5)
    api.Scene.from_flat(flat).build(api.BuildOptions(), listener)
    assert seen["rect"] == (2, 2, 6, 6)


def test_empty_paths_stay_out_of_a_dilated_batch():
    """An outline without contours has zero bounds; dilated they overlap the view box's corner. The CPU tiler builds
    one blank tile for such a path (never listed in a batch); the D3D11-level builder skips it like without dilation."""
    b = SceneBuilderPy((0, 0, 64, 64))
    b.end_path((1, 2, 3, 255))                                    # nothing pushed
    b.move_to(8, 8); b.line_to(40, 12); b.line_to(20, 44); b.close()
    b.end_path((200, 0, 0, 255))
    flat = b.finish("with-empty")
    for dilation in ((0.0, 0.0), (1.0, 1.0)):
        cmds = collect(api.Scene.from_flat(flat), api.BuildOptions(dilation=dilation))
        draw = [c for c in cmds if c["kind"] == "DrawTilesD3D11"][0]
        assert draw["path_count"] == 1 and draw["global_path_ids"] == [1]
        built = H.oracle_build(flat, None, dilation=dilation)
        assert len(built.tiles) > 0 and (built.tiles["path_id"] == 1).all()
