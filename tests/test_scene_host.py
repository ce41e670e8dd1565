"""Host side above the C ABI (csrc/scene.cpp): the D3D11-level SceneBuilder mirror, exercised
without a GPU through the RenderCommandListener callback."""
import ctypes as C

import numpy as np
import pytest

from pathfinder_b200 import _lib as L
from pathfinder_b200 import api, scenes
from pathfinder_b200.flat_scene import SceneBuilderPy
from tests import helpers as H


def collect(scene, options, sink=None):
    out = []

    def listener(cmd):
        rec = {"kind": L.COMMAND_NAMES[cmd.kind]}
        if cmd.kind == L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11:
            ds = cmd.u.upload_scene_d3d11.draw_segments
            rec["points"] = np.ctypeslib.as_array(C.cast(ds.points, C.POINTER(C.c_float)), (ds.point_count, 2)).copy()
            rec["indices"] = np.ctypeslib.as_array(C.cast(ds.indices, C.POINTER(C.c_uint32)), (ds.index_count, 2)).copy()
        elif cmd.kind == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            b = cmd.u.draw_tiles_d3d11.tile_batch_data
            rec.update(path_count=b.path_count, tile_count=b.tile_count, segment_count=b.segment_count,
                       content_key=b.content_key, batch_id=b.batch_id)
            t = b.prepare_info.transform
            rec["transform"] = (t.matrix.m00, t.matrix.m01, t.matrix.m10, t.matrix.m11, t.vector.x, t.vector.y)
            pm = C.cast(b.prepare_info.propagate_metadata, C.POINTER(L.PFPropagateMetadataD3D11))
            dm = C.cast(b.prepare_info.dice_metadata, C.POINTER(L.PFDiceMetadataD3D11))
            tp = C.cast(b.prepare_info.tile_path_info, C.POINTER(L.PFTilePathInfoD3D11))
            rec["rects"] = [(pm[i].tile_rect.origin.x, pm[i].tile_rect.origin.y, pm[i].tile_rect.lower_right.x,
                             pm[i].tile_rect.lower_right.y) for i in range(b.path_count)]
            rec["tile_offsets"] = [pm[i].tile_offset for i in range(b.path_count)]
            rec["z_write"] = [pm[i].z_write for i in range(b.path_count)]
            rec["global_path_ids"] = [dm[i].global_path_id for i in range(b.path_count)]
            rec["first_batch_segment"] = [dm[i].first_batch_segment_index for i in range(b.path_count)]
            rec["ctrl"] = [tp[i].ctrl for i in range(b.path_count)]
            rec["color"] = [tp[i].color for i in range(b.path_count)]
        elif cmd.kind == L.PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA:
            n = cmd.u.upload_texture_metadata.entry_count
            e = C.cast(cmd.u.upload_texture_metadata.entries, C.POINTER(L.PFTextureMetadataEntry))
            rec["colors"] = [(e[i].base_color.r, e[i].base_color.g, e[i].base_color.b, e[i].base_color.a) for i in range(n)]
        out.append(rec)

    scene.build(options, listener, sink)
    return out


def small_scene():
    b = SceneBuilderPy((0, 0, 64, 64))
    b.move_to(4, 4)
    b.line_to(40, 8)
    b.quad_to(50, 30, 30, 40)
    b.cubic_to(20, 50, 10, 45, 6, 30)
    b.close()
    b.end_path((255, 0, 0, 255))
    b.move_to(100, 100)  # outside the view box: skipped by prepare_draw_path_for_gpu_binning
    b.line_to(120, 100)
    b.line_to(110, 120)
    b.close()
    b.end_path((0, 255, 0, 128), fill_rule=1)
    b.move_to(20, 20)
    b.line_to(60, 20)
    b.line_to(60, 60)
    b.close()
    b.end_path((0, 255, 0, 128), fill_rule=1)
    return b.finish("small")


def test_command_order_and_payloads():
    flat = small_scene()
    cmds = collect(api.Scene.from_flat(flat), api.BuildOptions())
    # SceneBuilder::build order (renderer/src/builder.rs:160-221)
    assert [c["kind"] for c in cmds] == ["Start", "UploadTextureMetadata", "UploadSceneD3D11", "DrawTilesD3D11", "Finish"]
    assert cmds[1]["colors"] == [(255, 0, 0, 255), (0, 255, 0, 128)]  # Palette dedups equal paints
    up = cmds[2]
    # SegmentsD3D11::add_path (builder.rs:804-841): one index per on-curve point, first point re-appended.
    idx = up["indices"]
    assert idx[:4, 0].tolist() == [0, 1, 3, 6]
    assert idx[:4, 1].tolist() == [0, 0x80000000, 0x40000000, 0]
    pts = up["points"]
    assert np.array_equal(pts[7], pts[0])  # implicit close
    draw = cmds[3]
    assert draw["batch_id"] == 32  # MAX_CLIP_BATCHES
    assert draw["path_count"] == 2 and draw["global_path_ids"] == [0, 2]  # path 1 is off-screen
    assert draw["rects"][0] == (0, 0, 4, 4)  # bounds include control points (outline.rs:884-896)
    assert draw["rects"][1] == (1, 1, 4, 4)
    assert draw["tile_offsets"] == [0, 16] and draw["tile_count"] == 25
    assert draw["z_write"] == [1, 0]  # opaque paint occludes, translucent does not
    assert draw["ctrl"] == [1, 2] and draw["color"] == [0, 1]
    assert draw["first_batch_segment"] == [0, 4] and draw["segment_count"] == 7


def test_scene_upload_only_when_dirty():
    flat = small_scene()
    scene = api.Scene.from_flat(flat)
    sink = L.PFSceneSinkState()
    first = collect(scene, api.BuildOptions(), sink)
    second = collect(scene, api.BuildOptions(), sink)
    assert "UploadSceneD3D11" in [c["kind"] for c in first]
    assert "UploadSceneD3D11" not in [c["kind"] for c in second]  # builder.rs:193-215
    assert first[3]["content_key"] == second[2]["content_key"] != 0
    scene.set_view_box(flat.view_box)  # bumps the epoch
    third = collect(scene, api.BuildOptions(), sink)
    assert "UploadSceneD3D11" in [c["kind"] for c in third]
    assert third[3]["content_key"] != first[3]["content_key"]
    moved = collect(scene, api.BuildOptions(transform=api.Transform2F.from_scale(0.5)), sink)
    assert moved[2]["content_key"] != third[3]["content_key"]


@pytest.mark.parametrize("size", [256, 1024])
def test_tile_rects_match_cpu_tiler(size):
    """Sum of dense tile-map areas equals the CPU tiler's (renderer/src/builder.rs:430-436) for the
    scale + translate transform of the tiger."""
    flat, xf = scenes.tiger(size)
    cmds = collect(api.Scene.from_flat(flat), api.BuildOptions(transform=api.Transform2F(*xf)))
    draw = [c for c in cmds if c["kind"] == "DrawTilesD3D11"][0]
    built = H.oracle_build(flat, xf)
    assert draw["tile_count"] == built.bbox_tile_count
    assert draw["segment_count"] == built.input_segment_count


def test_random_scene_rects_match_cpu_tiler():
    flat = scenes.random_paths(2000, 2048, 7)
    cmds = collect(api.Scene.from_flat(flat), api.BuildOptions())
    draw = [c for c in cmds if c["kind"] == "DrawTilesD3D11"][0]
    built = H.oracle_build(flat, None)
    assert draw["tile_count"] == built.bbox_tile_count


def test_text_page_rects_match_cpu_tiler():
    """Many tiny paths of quadratics (BASELINE.json configs[2] outlines): same dense tile maps and segment count."""
    flat = scenes.text_page(2000, 1024, layout="lines")
    cmds = collect(api.Scene.from_flat(flat), api.BuildOptions())
    draw = [c for c in cmds if c["kind"] == "DrawTilesD3D11"][0]
    built = H.oracle_build(flat, None)
    assert draw["path_count"] == 2000
    assert draw["tile_count"] == built.bbox_tile_count
    assert draw["segment_count"] == built.input_segment_count


def test_unsupported_options_are_refused():
    flat = small_scene()
    with pytest.raises(L.PathfinderCudaError) as e:
        collect(api.Scene.from_flat(flat), api.BuildOptions(subpixel_aa_enabled=True))
    assert e.value.status == L.PF_CUDA_ERROR_UNSUPPORTED


def test_clip_batch_commands():
    """A scene with clipped draw paths emits PrepareClipTilesD3D11 before DrawTilesD3D11 (builder.rs:1098-1104);
    the draw batch's propagate metadata points into the clip batch and carries clipped_path_info."""
    from pathfinder_b200 import _lib as L
    from pathfinder_b200 import api
    from pathfinder_b200.flat_scene import SceneBuilderPy
    b = SceneBuilderPy((0, 0, 128, 128))
    b.move_to(10, 10); b.line_to(100, 20); b.line_to(50, 110); b.close()
    unused = b.end_clip_path()
    b.move_to(20, 20); b.line_to(120, 20); b.line_to(120, 120); b.line_to(20, 120); b.close()
    used = b.end_clip_path()
    b.move_to(0, 0); b.line_to(64, 0); b.line_to(64, 64); b.close()
    b.end_path((255, 0, 0, 255))
    b.move_to(0, 0); b.line_to(128, 0); b.line_to(128, 128); b.line_to(0, 128); b.close()
    b.end_path((0, 255, 0, 255), clip=used)
    flat = b.finish()
    assert unused == 0 and used == 1
    scene = api.Scene.from_flat(flat)
    seen = []

    def listener(cmd):
        kind = int(cmd.kind)
        seen.append(kind)
        if kind == L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11:
            u = cmd.u.upload_scene_d3d11
            assert u.clip_segments.index_count == 3 + 4 and u.draw_segments.index_count == 3 + 4
        if kind == L.PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11:
            batch = cmd.u.prepare_clip_tiles_d3d11.batch
            assert batch.path_count == 1 and batch.path_source == 1
            pm = C.cast(batch.prepare_info.propagate_metadata, C.POINTER(L.PFPropagateMetadataD3D11))[0]
            dm = C.cast(batch.prepare_info.dice_metadata, C.POINTER(L.PFDiceMetadataD3D11))
            assert (pm.tile_rect.origin.x, pm.tile_rect.origin.y, pm.tile_rect.lower_right.x, pm.tile_rect.lower_right.y) == (1, 1, 8, 8)
            assert batch.tile_count == 49 and batch.segment_count == 4
            assert dm[0].global_path_id == used
            assert dm[0].first_global_segment_index == 3  # after the unused clip path
        if kind == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            batch = cmd.u.draw_tiles_d3d11.tile_batch_data
            assert batch.has_clipped_path_info == 1
            assert batch.clipped_path_info.clipped_path_count == 1 and batch.clipped_path_info.max_clipped_tile_count == 64
            pms = C.cast(batch.prepare_info.propagate_metadata, C.POINTER(L.PFPropagateMetadataD3D11))
            assert pms[0].clip_path_index == 0xFFFFFFFF and pms[1].clip_path_index == 0

    scene.build(api.BuildOptions(), listener)
    i_clip = seen.index(L.PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11)
    assert seen.index(L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11) < i_clip < seen.index(L.PF_RENDER_COMMAND_DRAW_TILES_D3D11)


def test_concurrent_scene_builds_share_the_worker_pool():
    """Scene::build from several host threads at once (ctypes releases the GIL): the builds share one worker pool,
    which runs one parallel region at a time, and every build must produce the arrays a lone build produces."""
    import threading
    flat = scenes.random_paths(30000, 2048, 77)

    def build_once():
        scene = api.Scene.from_flat(flat)
        out = {}

        def listener(cmd):
            if int(cmd.kind) == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
                b = cmd.u.draw_tiles_d3d11.tile_batch_data
                n = int(b.path_count)
                pm = np.ctypeslib.as_array(C.cast(b.prepare_info.propagate_metadata, C.POINTER(C.c_uint8)),
                                           shape=(n * C.sizeof(L.PFPropagateMetadataD3D11),)).copy()
                dm = np.ctypeslib.as_array(C.cast(b.prepare_info.dice_metadata, C.POINTER(C.c_uint8)),
                                           shape=(n * C.sizeof(L.PFDiceMetadataD3D11),)).copy()
                out["batch"] = (n, int(b.tile_count), int(b.segment_count), pm.tobytes(), dm.tobytes())
            if int(cmd.kind) == L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11:
                s = cmd.u.upload_scene_d3d11.draw_segments
                pts = np.ctypeslib.as_array(C.cast(s.points, C.POINTER(C.c_float)), shape=(int(s.point_count) * 2,)).copy()
                out["segments"] = (int(s.point_count), int(s.index_count), pts.tobytes())

        for _ in range(3):
            scene.set_view_box(flat.view_box)  # dirty: segments and batch are rebuilt
            scene.build(api.BuildOptions(), listener)
        return out

    expected = build_once()
    results, errors = [None] * 6, []

    def worker(i):
        try:
            results[i] = build_once()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(results))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for r in results:
        assert r == expected


def test_push_calls_refuse_arguments_they_could_not_index():
    """An unknown paint id, fill rule or a decreasing offset table is refused when pushed (the reference panics
    later, when it indexes the palette or the outline); a refused call leaves the scene as it was."""
    lib = L.lib()
    scene = api.Scene()
    scene.set_view_box((0, 0, 64, 64))
    paint = scene.push_paint((255, 0, 0, 255))
    tri = np.array([(1, 1), (30, 2), (10, 40)], np.float32)
    ok = scene.push_draw_path(tri, [0, 0, 0], [0, 3], paint)
    assert ok == 0 and scene.draw_path_count() == 1
    epoch = lib.PFSceneGetEpoch(scene._h)
    none = 0xFFFFFFFF
    assert scene.push_draw_path(tri, [0, 0, 0], [0, 3], paint + 1) == none
    assert b"paint" in lib.PFCudaGetLastError()
    assert scene.push_draw_path(tri, [0, 0, 0], [0, 3], paint, fill_rule=7) == none
    assert scene.push_draw_path(tri, [0, 0, 0], [0, 3, 2], paint) == none
    assert scene.push_clip_path(tri, [0, 0, 0], [0, 3], fill_rule=9) == none
    assert scene.draw_path_count() == 1 and lib.PFSceneGetEpoch(scene._h) == epoch

    def push_many(contour_offsets, path_contour_offsets, paints, rules):
        pts = np.concatenate([tri, tri + 5]).astype(np.float32)
        fl = np.zeros(6, np.uint8)
        co = np.asarray(contour_offsets, np.uint32)
        pco = np.asarray(path_contour_offsets, np.uint32)
        pa, ru = np.asarray(paints, np.uint16), np.asarray(rules, np.uint8)
        return lib.PFScenePushDrawPaths(scene._h, pts.ctypes.data, fl.ctypes.data, 6, co.ctypes.data, len(co) - 1,
                                        pco.ctypes.data, len(pco) - 1, pa.ctypes.data, ru.ctypes.data, None)

    bad = L.PF_CUDA_ERROR_INVALID_ARGUMENT
    assert push_many([0, 3, 5], [0, 1, 2], [paint, paint], [0, 0]) == bad      # offsets do not cover the points
    assert push_many([0, 3, 6], [0, 1, 3], [paint, paint], [0, 0]) == bad      # path 1 reaches past the contours
    assert push_many([0, 3, 6], [0, 1, 2], [paint, paint + 3], [0, 0]) == bad  # second path's paint: nothing is kept
    assert push_many([0, 3, 6], [0, 1, 2], [paint, paint], [0, 2]) == bad
    assert scene.draw_path_count() == 1 and lib.PFSceneGetEpoch(scene._h) == epoch
    assert push_many([0, 3, 6], [0, 1, 2], [paint, paint], [0, 1]) == 0
    assert scene.draw_path_count() == 3


def test_display_list_with_a_render_target_becomes_the_reference_command_order():
    """Scene::push_render_target / pop_render_target + a pattern paint (scene.rs:110-123, paint.rs:138-146): the
    build emits AllocateTexturePage and DeclareRenderTarget before the metadata, one DrawTilesD3D11 per display item
    between PushRenderTarget / PopRenderTarget, and the batch that samples the target names its page
    (builder.rs:327-357, paint.rs:399-437,626-659)."""
    from pathfinder_b200 import scenes
    size = 128
    wide = scenes.text_page_subpixel(40, size, layout="lines")
    scene = scenes.subpixel_scene(wide, size)
    kinds, draws, paints = [], [], []

    def listener(cmd_ptr, _userdata):
        cmd = cmd_ptr.contents
        kinds.append(cmd.kind)
        if cmd.kind == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            d = cmd.u.draw_tiles_d3d11
            draws.append((d.tile_batch_data.path_count, d.has_color_texture, d.color_texture.page, d.tile_batch_data.batch_id))
        elif cmd.kind == L.PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA:
            e = C.cast(cmd.u.upload_texture_metadata.entries, C.POINTER(L.PFTextureMetadataEntry))
            for i in range(cmd.u.upload_texture_metadata.entry_count):
                t = e[i].color_0_transform
                paints.append((e[i].color_0_combine_mode, e[i].filter.kind, e[i].filter.flags,
                               (t.matrix.m00, t.matrix.m01, t.matrix.m10, t.matrix.m11, t.vector.x, t.vector.y)))
        elif cmd.kind == L.PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE:
            assert (cmd.u.allocate_texture_page.size.x, cmd.u.allocate_texture_page.size.y) == (3 * size, size)
        elif cmd.kind == L.PF_RENDER_COMMAND_DECLARE_RENDER_TARGET:
            r = cmd.u.declare_render_target.location.rect
            assert (r.origin.x, r.origin.y, r.lower_right.x, r.lower_right.y) == (0, 0, 3 * size, size)
        return 0

    sink = L.PFSceneSinkState(0, 0, 0)
    fn = L.LISTENER_FN(listener)
    assert L.lib().PFSceneBuild(scene._h, api.BuildOptions()._h, C.byref(sink), fn, None) == 0
    K = L
    assert kinds == [K.PF_RENDER_COMMAND_START, K.PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE, K.PF_RENDER_COMMAND_DECLARE_RENDER_TARGET,
                     K.PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA, K.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11,
                     K.PF_RENDER_COMMAND_PUSH_RENDER_TARGET, K.PF_RENDER_COMMAND_DRAW_TILES_D3D11,
                     K.PF_RENDER_COMMAND_POP_RENDER_TARGET, K.PF_RENDER_COMMAND_DRAW_TILES_D3D11, K.PF_RENDER_COMMAND_FINISH]
    assert draws[0][1] == 0 and draws[0][0] > 30                      # the glyphs, no colour texture
    assert draws[1][:3] == (1, 1, 0) and draws[1][3] == draws[0][3] + 1  # the page rectangle samples page 0
    solid, pattern = paints
    assert solid[:2] == (L.PF_COLOR_COMBINE_MODE_NONE, L.PF_FILTER_NONE)
    assert pattern[0] == L.PF_COLOR_COMBINE_MODE_SRC_IN and pattern[1] == L.PF_FILTER_TEXT
    assert pattern[2] == (L.PF_FILTER_FLAG_TEXT_HAS_KERNEL | L.PF_FILTER_FLAG_TEXT_GAMMA_CORRECTION)
    # pixel centre (x + 0.5, y + 0.5) of the page -> u = (3x + 1.5) / 3W, v = 1 - (y + 0.5) / H (bottom-up, like the GL backend)
    m00, m01, m10, m11, tx, ty = pattern[3]
    assert np.allclose([m00, m01, m10, m11, tx, ty], [1.0 / size, 0.0, 0.0, -1.0 / size, 0.0, 1.0], atol=1e-7)
    # a pop without a push is refused
    bad = api.Scene()
    bad.pop_render_target()
    assert L.lib().PFSceneBuild(bad._h, api.BuildOptions()._h, C.byref(L.PFSceneSinkState(0, 0, 0)), fn, None) == L.PF_CUDA_ERROR_PROTOCOL


def collect_strip(scene, options, strip):
    """collect() for a renderer that owns the tile rows `strip` (PFSceneBuildForStrip)."""
    out = []

    def listener(cmd):
        rec = {"kind": L.COMMAND_NAMES[cmd.kind]}
        if cmd.kind == L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11:
            ds = cmd.u.upload_scene_d3d11.draw_segments
            rec["point_count"], rec["index_count"] = int(ds.point_count), int(ds.index_count)
            rec["indices"] = np.ctypeslib.as_array(C.cast(ds.indices, C.POINTER(C.c_uint32)), (ds.index_count, 2)).copy()
            rec["points"] = np.ctypeslib.as_array(C.cast(ds.points, C.POINTER(C.c_float)), (ds.point_count, 2)).copy()
        elif cmd.kind == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            b = cmd.u.draw_tiles_d3d11.tile_batch_data
            pm = C.cast(b.prepare_info.propagate_metadata, C.POINTER(L.PFPropagateMetadataD3D11))
            dm = C.cast(b.prepare_info.dice_metadata, C.POINTER(L.PFDiceMetadataD3D11))
            rec.update(path_count=int(b.path_count), segment_count=int(b.segment_count), content_key=int(b.content_key))
            rec["rects"] = [(pm[i].tile_rect.origin.x, pm[i].tile_rect.origin.y, pm[i].tile_rect.lower_right.x,
                             pm[i].tile_rect.lower_right.y) for i in range(b.path_count)]
            rec["global_path_ids"] = [int(dm[i].global_path_id) for i in range(b.path_count)]
            rec["first_global_segment"] = [int(dm[i].first_global_segment_index) for i in range(b.path_count)]
            rec["first_batch_segment"] = [int(dm[i].first_batch_segment_index) for i in range(b.path_count)]
        out.append(rec)

    scene.build(options, listener, None, strip=strip)
    return out


@pytest.mark.parametrize("mode", ["transformed", "plain", "prepared"])
def test_build_for_a_strip_keeps_only_the_paths_that_reach_it(mode):
    """Multi-GPU host side: a rank builds segments and records only for the paths with a tile in its rows; ids stay
    global; the strips together cover exactly the paths of the whole build; the kept paths' segments are the same
    points in the same order. The three ways the builder gets a path's bounds: under the options' transform, as pushed
    (identity), and from the prepared (transformed + dilated) copy of the points."""
    flat = scenes.random_paths(600, 512, 21, r_min=4.0, r_max=40.0)
    options = {"transformed": lambda: api.BuildOptions(transform=api.Transform2F(0.9, 0.1, -0.05, 1.1, 6.0, -3.0)),
               "plain": lambda: api.BuildOptions(),
               "prepared": lambda: api.BuildOptions(transform=api.Transform2F(0.9, 0.1, -0.05, 1.1, 6.0, -3.0),
                                                    dilation=(0.5, 0.25))}[mode]()
    whole = collect_strip(api.Scene.from_flat(flat), options, None)
    whole_draw = next(r for r in whole if r["kind"] == "DrawTilesD3D11")
    whole_up = next(r for r in whole if r["kind"] == "UploadSceneD3D11")
    rect_of = dict(zip(whole_draw["global_path_ids"], whole_draw["rects"]))
    rows = 512 // 16
    seen = set()
    keys = set()
    for y0, y1 in ((0, 11), (11, 22), (22, rows)):
        scene = api.Scene.from_flat(flat)
        got = collect_strip(scene, options, (y0, y1))
        draw = next(r for r in got if r["kind"] == "DrawTilesD3D11")
        up = next(r for r in got if r["kind"] == "UploadSceneD3D11")
        want = [pid for pid, r in rect_of.items() if r[1] < y1 and r[3] > y0]
        assert draw["global_path_ids"] == want
        assert draw["rects"] == [rect_of[pid] for pid in want]  # full rects: the renderer restricts them to its rows
        # only the kept paths' segments were copied, contiguously and in path order
        assert up["index_count"] == draw["segment_count"] < len(whole_up["indices"])
        assert draw["first_batch_segment"] == draw["first_global_segment"]
        # ... and they are the whole build's segments of those paths: same points behind the same flags
        w_first = dict(zip(whole_draw["global_path_ids"], whole_draw["first_global_segment"]))
        w_order = whole_draw["global_path_ids"]
        for k, pid in enumerate(want[:40]):
            a0 = draw["first_global_segment"][k]
            a1 = draw["first_global_segment"][k + 1] if k + 1 < len(want) else up["index_count"]
            j = w_order.index(pid)
            b0 = w_first[pid]
            # (the whole build's upload also holds the segments of paths outside the view box: count from the batch)
            wb = whole_draw["first_batch_segment"]
            b1 = b0 + ((wb[j + 1] if j + 1 < len(w_order) else whole_draw["segment_count"]) - wb[j])
            assert a1 - a0 == b1 - b0
            assert np.array_equal(up["indices"][a0:a1, 1], whole_up["indices"][b0:b1, 1])
            assert np.array_equal(up["points"][up["indices"][a0:a1, 0]], whole_up["points"][whole_up["indices"][b0:b1, 0]])
        seen.update(want)
        keys.add(draw["content_key"])
        # the same strip again: nothing is rebuilt or re-sent (same content key); another strip re-uploads
        sink = L.PFSceneSinkState()
        again = []
        scene.build(options, lambda c: again.append(L.COMMAND_NAMES[c.kind]), sink, strip=(y0, y1))
        again2 = []
        scene.build(options, lambda c: again2.append(L.COMMAND_NAMES[c.kind]), sink, strip=(y0, y1))
        assert "UploadSceneD3D11" in again and "UploadSceneD3D11" not in again2
        again3 = []
        scene.build(options, lambda c: again3.append(L.COMMAND_NAMES[c.kind]), sink, strip=(y0, y1 - 1))
        assert "UploadSceneD3D11" in again3
    assert seen == set(whole_draw["global_path_ids"])
    assert len(keys) == 3 and whole_draw["content_key"] not in keys


def test_strip_inclusion_edge_cases_and_chunks():
    """The strip filter is exactly "the whole build's tile rect has a row in the strip": bounds that end on a strip
    boundary, paths off every side of the view box, a NaN point, a frame-sized path; an empty strip (= no strip) and a
    strip below the frame; enough paths that the inclusion pass runs in several chunks on the worker pool."""
    flat = scenes.random_paths(20000, 1024, 77, r_min=2.0, r_max=30.0)
    scene_whole, scene_strip = api.Scene.from_flat(flat), api.Scene.from_flat(flat)

    def box(x0, y0, x1, y1):
        return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], dtype=np.float32)

    extra = [box(10, 100, 50, 320),          # max_y on the boundary of rows 20: not in a strip that starts there
             box(10, 320, 50, 400),          # min_y on it: not in a strip that ends there
             box(10, 319.5, 50, 320.5),      # straddles it
             box(-500, 300, -20, 340),       # left of the view box
             box(1500, 300, 1600, 340),      # right of it
             box(100, -300, 140, -20),       # above
             box(100, 1100, 140, 1300),      # below
             box(-100, -100, 1200, 1200),    # larger than the frame
             np.array([[5, 5], [float("nan"), 50], [60, 60], [5, 60]], dtype=np.float32),
             box(200, 0, 240, 1024)]         # every row
    for s in (scene_whole, scene_strip):
        for pts in extra:
            s.push_draw_path(pts, np.zeros(len(pts), dtype=np.uint8), np.array([0, len(pts)], dtype=np.uint32), 0)
    options = api.BuildOptions()
    whole = next(r for r in collect_strip(scene_whole, options, None) if r["kind"] == "DrawTilesD3D11")
    rect_of = dict(zip(whole["global_path_ids"], whole["rects"]))
    for y0, y1 in ((0, 20), (20, 21), (20, 64), (63, 64), (30, 30), (64, 80), (0, 64)):
        got = [r for r in collect_strip(scene_strip, options, (y0, y1)) if r["kind"] == "DrawTilesD3D11"]
        want = [pid for pid, r in rect_of.items() if r[1] < y1 and r[3] > y0]
        if y0 == y1:  # no rows: not a strip, the scene is built whole (a renderer never owns an empty strip)
            want = list(rect_of)
        if not want:
            assert not got or got[0]["global_path_ids"] == []
            continue
        assert got[0]["global_path_ids"] == want, (y0, y1)
        assert got[0]["rects"] == [rect_of[pid] for pid in want]


@pytest.mark.parametrize("strip", [None, (0, 4)])
def test_empty_and_ragged_scenes(strip):
    """No paths, a path without contours, a contour without points: the builder still sends a well-formed frame
    (Start .. Finish) and no batch, for a whole build and for a strip; the zero-area bounds of an empty outline do not
    intersect the view box (RectF::intersection is strict, rect.rs:122-137), so prepare_draw_path_for_gpu_binning
    (builder.rs:1075-1079) drops the path. A real path pushed after them keeps its global id."""
    frame = ["Start", "UploadTextureMetadata", "UploadSceneD3D11", "Finish"]
    scene = api.Scene()
    scene.set_view_box((0, 0, 256, 256))
    assert [r["kind"] for r in collect_strip(scene, api.BuildOptions(), strip)] == frame
    scene.push_paint((255, 0, 0, 255))
    none = np.zeros((0, 2), np.float32), np.zeros(0, np.uint8)
    scene.push_draw_path(*none, np.array([0], np.uint32), 0)
    assert [r["kind"] for r in collect_strip(scene, api.BuildOptions(), strip)] == frame
    scene.push_draw_path(*none, np.array([0, 0], np.uint32), 0)
    assert [r["kind"] for r in collect_strip(scene, api.BuildOptions(), strip)] == frame
    scene.push_draw_path(np.array([[10, 10], [50, 10], [50, 50]], np.float32), np.zeros(3, np.uint8), np.array([0, 3], np.uint32), 0)
    got = collect_strip(scene, api.BuildOptions(), strip)
    assert [r["kind"] for r in got] == frame[:3] + ["DrawTilesD3D11", "Finish"]
    draw = got[3]
    assert draw["path_count"] == 1 and draw["global_path_ids"] == [2] and draw["rects"] == [(0, 0, 4, 4)]
    assert draw["segment_count"] == 3 == got[2]["index_count"]
