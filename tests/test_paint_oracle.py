"""SURVEY.md §8 f4 on the CPU: the paint oracle (oracle/paint_oracle.py) against the one known-answer test the
reference holds near it and against hand-computed values, and the product's host side (gradient ramps, texture
pages, texture transforms, blend-mode entries as PFSceneBuild emits them) against the oracle. No GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import paint_oracle as P
from pathfinder_b200 import _lib as L
from pathfinder_b200 import api

RECT = np.asarray([[8, 8], [56, 8], [56, 56], [8, 56]], np.float32)


def test_reference_known_answer_never_sample_zero_width():
    """content/src/gradient.rs:291-303, restated: 110 stops at offsets (i % 11) / 10, the zero-width ones red."""
    stops = []
    for i in range(110):
        zero_width = i == 0 or 11 <= i < 99 or i == 109
        stops.append((np.float32(i % 11) / np.float32(10.0), (255 if zero_width else 0, 0, 0, 1)))
    stops = P.sort_stops(stops)
    # stable_order (gradient.rs:278-289): sorted by offset, insertion order kept among equals
    assert all(a[0] <= b[0] for a, b in zip(stops, stops[1:]))
    for i in range(11):
        assert P.gradient_sample(stops, np.float32(i) / np.float32(10.0))[0] == 0, i


def test_gradient_sample_by_hand():
    stops = [(0.0, (255, 0, 0, 255)), (0.5, (0, 255, 0, 255)), (1.0, (0, 0, 255, 255))]
    assert P.gradient_sample(stops, 0.0) == (255, 0, 0, 255)
    assert P.gradient_sample(stops, 0.25) == (128, 128, 0, 255)  # 127.5 rounds to even
    assert P.gradient_sample(stops, 0.5) == (0, 255, 0, 255)
    assert P.gradient_sample(stops, 2.0) == (0, 0, 255, 255)  # clamped
    assert P.gradient_sample([], 0.3) == (0, 0, 0, 0)
    assert P.gradient_sample([(0.3, (9, 8, 7, 6))], 0.9) == (9, 8, 7, 6)


def test_blend_modes_by_hand():
    d = np.zeros((1, 1, 4), np.float32)
    d[0, 0] = (0.5, 0.25, 0.75, 1.0)
    c, one = (0.2, 0.6, 1.0, 1.0), np.ones((1, 1), np.float32)
    near = lambda got, want: np.allclose(got[0, 0], want, atol=1e-6)  # noqa: E731
    assert near(P.blend(d, c, one, "src_over"), (0.2, 0.6, 1.0, 1.0))
    assert near(P.blend(d, c, one * 0.5, "src_over"), (0.35, 0.425, 0.875, 1.0))
    assert near(P.blend(d, c, one, "multiply"), (0.1, 0.15, 0.75, 1.0))
    assert near(P.blend(d, c, one, "screen"), (0.6, 0.7, 1.0, 1.0))
    assert near(P.blend(d, c, one, "darken"), (0.2, 0.25, 0.75, 1.0))
    assert near(P.blend(d, c, one, "lighten"), (0.5, 0.6, 1.0, 1.0))
    assert near(P.blend(d, c, one, "difference"), (0.3, 0.35, 0.25, 1.0))
    assert near(P.blend(d, c, one, "exclusion"), (0.5, 0.55, 0.25, 1.0))
    assert near(P.blend(d, c, one, "dest_over"), d[0, 0])            # opaque destination: nothing shows through
    assert near(P.blend(d, c, one, "dest_out"), (0.0, 0.0, 0.0, 0.0))
    assert near(P.blend(d, c, one, "lighter"), (0.7, 0.85, 1.0, 1.0))  # clamped
    assert near(P.blend(d, c, one * 0.0, "hue"), d[0, 0])            # no coverage: the pixel is kept (alpha forced to 1)
    # luminosity of a grey source over a colour keeps hue and saturation: HSL round trip of the destination
    grey = (0.5, 0.5, 0.5, 1.0)
    lum = P.blend(d, grey, one, "color")[0, 0]
    assert np.allclose(lum[:3], 0.5, atol=1e-6)                      # colour of a grey source = grey at the dest's lightness
    with pytest.raises(ValueError):
        P.blend(d, c, one, "copy")


def test_sampler_by_hand():
    tex = np.zeros((2, 4, 4), np.uint8)
    tex[0, :, 0] = (0, 85, 170, 255)
    tex[1, :, 0] = 255
    u = np.asarray([0.125, 0.25, 0.875, 1.2], np.float32)  # texel centre 0, between 0 and 1, centre 3, beyond
    v = np.full(4, 0.25, np.float32)                       # centre of row 0
    got = P.sample(tex, u, v)[:, 0]
    assert np.allclose(got, [0.0, (85 / 255) * 0.5, 1.0, 1.0], atol=1e-6)
    assert np.allclose(P.sample(tex, u, v, P.REPEAT_U)[3, 0], (0.8 * 0 + 0.2 * 85) / 255 * 0 + P.sample(tex, np.float32([0.2]), v[:1])[0, 0], atol=1e-6)
    assert np.allclose(P.sample(tex, u, v, P.NEAREST)[:, 0], [0.0, 85 / 255, 1.0, 1.0], atol=1e-6)
    # bottom-up addressing (render targets): v = 0.75 is the centre of the TOP row
    assert np.allclose(P.sample(tex, u[:1], np.float32([0.75]), bottom_up=True)[0, 0], 0.0)


def collect(scene, options=None):
    """Runs PFSceneBuild and keeps what the paint path sends: pages, texel uploads, texture metadata, batches."""
    out = {"pages": {}, "uploads": [], "meta": [], "batches": [], "start": None}

    def listener(cmd):
        if cmd.kind == L.PF_RENDER_COMMAND_START:
            out["start"] = int(cmd.u.start.needs_readable_framebuffer)
        elif cmd.kind == L.PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE:
            a = cmd.u.allocate_texture_page
            out["pages"][int(a.page_id)] = (int(a.size.x), int(a.size.y))
        elif cmd.kind == L.PF_RENDER_COMMAND_UPLOAD_TEXEL_DATA:
            up = cmd.u.upload_texel_data
            rect = up.location.rect
            w, h = rect.lower_right.x - rect.origin.x, rect.lower_right.y - rect.origin.y
            assert up.texel_count == w * h
            texels = np.ctypeslib.as_array(C.cast(up.texels, C.POINTER(C.c_uint8)), shape=(h, w, 4)).copy()
            out["uploads"].append((int(up.location.page), (rect.origin.x, rect.origin.y), texels))
        elif cmd.kind == L.PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA:
            m = cmd.u.upload_texture_metadata
            entries = C.cast(m.entries, C.POINTER(L.PFTextureMetadataEntry))
            for i in range(m.entry_count):
                e = entries[i]
                t = e.color_0_transform
                out["meta"].append({"transform": (t.matrix.m00, t.matrix.m01, t.matrix.m10, t.matrix.m11, t.vector.x, t.vector.y),
                                    "combine": int(e.color_0_combine_mode), "blend": int(e.blend_mode),
                                    "filter": int(e.filter.kind), "params": [float(x) for x in e.filter.params],
                                    "base": (e.base_color.r, e.base_color.g, e.base_color.b, e.base_color.a)})
        elif cmd.kind == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            d = cmd.u.draw_tiles_d3d11
            infos = C.cast(d.tile_batch_data.prepare_info.tile_path_info, C.POINTER(L.PFTilePathInfoD3D11))
            props = C.cast(d.tile_batch_data.prepare_info.propagate_metadata, C.POINTER(L.PFPropagateMetadataD3D11))
            n = d.tile_batch_data.path_count
            out["batches"].append({"texture": (int(d.color_texture.page), int(d.color_texture.sampling_flags)) if d.has_color_texture else None,
                                   "colors": [int(infos[i].color) for i in range(n)],
                                   "z_write": [int(props[i].z_write) for i in range(n)],
                                   "tile_rects": [(props[i].tile_rect.origin.x, props[i].tile_rect.origin.y,
                                                   props[i].tile_rect.lower_right.x, props[i].tile_rect.lower_right.y)
                                                  for i in range(n)]})

    scene.build(options or api.BuildOptions(), listener)
    return out


def push_rect(scene, paint, blend="src_over"):
    return scene.push_draw_path(RECT, np.zeros(4, np.uint8), [0, 4], paint, blend_mode=api.BLEND_MODES[blend])


def test_gradient_ramp_and_transform_as_built():
    rng = np.random.default_rng(7)
    scene = api.Scene()
    scene.set_view_box((0, 0, 64, 64))
    gradients = []
    for k in range(3):
        stops = sorted((float(np.float32(o)), tuple(int(c) for c in rng.integers(0, 256, 4))) for o in rng.random(5))
        line = ((3.0 + k, 5.0), (60.0, 40.0 + k))
        gradients.append((stops, line, scene.push_gradient(stops, line, repeat=(k == 1))))
    radial = scene.push_gradient([(0.0, (255, 255, 255, 255)), (1.0, (0, 0, 0, 255))], ((32, 32), (40, 36)), radii=(2.0, 30.0),
                                 transform=(2.0, 0.0, 0.0, 0.5, 1.0, -3.0))
    for _, _, pid in gradients:
        push_rect(scene, pid)
    push_rect(scene, radial)
    xf = api.Transform2F(0.5, 0.0, 0.0, 2.0, 3.0, -1.0)
    got = collect(scene, api.BuildOptions(transform=xf))
    assert got["pages"] == {0: (256, 256)}
    (page, origin, texels), = got["uploads"]
    assert page == 0 and origin == (0, 0)
    want = P.gradient_page([P.gradient_ramp(stops) for stops, _, _ in gradients] +
                           [P.gradient_ramp([(0.0, (255, 255, 255, 255)), (1.0, (0, 0, 0, 255))])])
    assert np.array_equal(texels, want)
    render_transform = P.t_inverse((0.5, 0.0, 0.0, 2.0, 3.0, -1.0))
    for row, (stops, line, pid) in enumerate(gradients):
        e = got["meta"][pid]
        assert e["combine"] == L.PF_COLOR_COMBINE_MODE_SRC_IN and e["filter"] == L.PF_FILTER_NONE and e["base"] == (255, 255, 255, 255)
        assert np.allclose(e["transform"], P.linear_gradient_transform(line, row, render_transform), rtol=1e-5, atol=1e-7)
    e = got["meta"][radial]
    t, p0, p1 = P.radial_gradient_entry(((32, 32), (40, 36)), (2.0, 30.0), 3, (2.0, 0.0, 0.0, 0.5, 1.0, -3.0), render_transform)
    assert e["filter"] == L.PF_FILTER_RADIAL_GRADIENT
    assert np.allclose(e["transform"], t, rtol=1e-5, atol=1e-7)
    assert np.allclose(e["params"][:8], [32, 32, 40, 36, 2.0, 30.0, 0.0, 3.5 / 256.0])
    # one batch per sampler state: the repeating gradient breaks the batch twice
    assert [b["texture"] for b in got["batches"]] == [(0, 0), (0, L.PF_TEXTURE_SAMPLING_FLAGS_REPEAT_U), (0, 0)]


def test_image_pattern_pages_as_built():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (5, 7, 4), dtype=np.uint8)
    scene = api.Scene()
    scene.set_view_box((0, 0, 64, 64))
    a = scene.push_image_pattern(img, transform=(2.0, 0.0, 0.0, 3.0, 10.0, 4.0))
    b = scene.push_image_pattern(img, repeat_x=True, repeat_y=True, smoothing=False)
    push_rect(scene, a)
    push_rect(scene, b)
    got = collect(scene)
    assert got["pages"] == {0: (9, 7), 1: (7, 5)}  # a texel of border only where the image does not repeat
    assert [(p, o) for p, o, _ in got["uploads"]] == [(0, (1, 1)), (1, (0, 0))]
    assert all(np.array_equal(t, img) for _, _, t in got["uploads"])
    assert np.allclose(got["meta"][a]["transform"], P.image_pattern_transform((9, 7), (2.0, 0.0, 0.0, 3.0, 10.0, 4.0)), rtol=1e-6)
    assert np.allclose(got["meta"][b]["transform"], P.image_pattern_transform((7, 5)), rtol=1e-6)
    nearest = L.PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MIN | L.PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MAG
    repeat = L.PF_TEXTURE_SAMPLING_FLAGS_REPEAT_U | L.PF_TEXTURE_SAMPLING_FLAGS_REPEAT_V
    assert [x["texture"] for x in got["batches"]] == [(0, 0), (1, repeat | nearest)]


def test_blend_modes_get_their_own_metadata_entries():
    scene = api.Scene()
    scene.set_view_box((0, 0, 64, 64))
    red, blue = scene.push_paint((255, 0, 0, 255)), scene.push_paint((0, 0, 255, 255))
    push_rect(scene, red)
    push_rect(scene, blue, "multiply")
    push_rect(scene, blue, "multiply")
    push_rect(scene, red, "xor")
    push_rect(scene, blue)
    got = collect(scene)
    assert got["start"] == 1  # Multiply reads the destination (BlendModeExt::needs_readable_framebuffer)
    assert [m["blend"] for m in got["meta"]] == [api.BLEND_MODES["src_over"]] * 2 + [api.BLEND_MODES["multiply"], api.BLEND_MODES["xor"]]
    assert got["meta"][2]["base"] == (0, 0, 255, 255) and got["meta"][3]["base"] == (255, 0, 0, 255)
    (batch,) = got["batches"]
    assert batch["colors"] == [red, 2, 2, 3, blue] and batch["texture"] is None
    assert batch["z_write"] == [1, 0, 0, 0, 1]  # occludes = opaque && SrcOver (builder.rs:83)
    assert batch["tile_rects"] == [(0, 0, 4, 4)] * 5  # RECT rounded out to tiles


def test_destructive_blend_modes_are_tiled_over_the_view_box():
    """BuiltPath::new (builder.rs:430-434): tile map bounds = the view box for Clear, Copy, SrcIn, DestIn, SrcOut and
    DestAtop; an opaque Clear occludes like an opaque SrcOver (effects.rs:202-204, builder.rs:83)."""
    scene = api.Scene()
    scene.set_view_box((0, 0, 100, 70))
    opaque, translucent = scene.push_paint((1, 2, 3, 255)), scene.push_paint((1, 2, 3, 128))
    small = RECT * 0.25 + 20  # 22..34: tiles (1, 1) .. (3, 3)
    modes = ["clear", "copy", "src_in", "dest_in", "src_out", "dest_atop", "clear", "src_atop"]
    for k, mode in enumerate(modes):
        scene.push_draw_path(small, np.zeros(4, np.uint8), [0, 4], translucent if k == 6 else opaque, blend_mode=api.BLEND_MODES[mode])
    got = collect(scene)
    (batch,) = got["batches"]
    assert batch["tile_rects"] == [(0, 0, 7, 5)] * 7 + [(1, 1, 3, 3)]
    assert batch["z_write"] == [1, 0, 0, 0, 0, 0, 0, 0]
    assert [got["meta"][c]["blend"] for c in batch["colors"]] == [api.BLEND_MODES[m] for m in modes]


def test_clip_paths_in_a_display_list_build():
    """A scene that takes the display-list builder (a gradient) and has clipped paths: one PrepareClipTilesD3D11 before
    the first draw batch that needs it, clip indices in the batches' records."""
    scene = api.Scene()
    scene.set_view_box((0, 0, 64, 64))
    red = scene.push_paint((255, 0, 0, 255))
    g = scene.push_gradient([(0.0, (255, 0, 0, 255)), (1.0, (0, 0, 255, 255))], ((0, 0), (64, 0)))
    clip = scene.push_clip_path(RECT * 0.5 + 10, np.zeros(4, np.uint8), [0, 4])
    push_rect(scene, red)
    scene.push_draw_path(RECT, np.zeros(4, np.uint8), [0, 4], g, clip_path_id=clip)
    scene.push_draw_path(RECT, np.zeros(4, np.uint8), [0, 4], red, clip_path_id=clip)
    kinds, clipped = [], []

    def listener(cmd):
        kinds.append(L.COMMAND_NAMES[cmd.kind])
        if cmd.kind == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            b = cmd.u.draw_tiles_d3d11.tile_batch_data
            pm = C.cast(b.prepare_info.propagate_metadata, C.POINTER(L.PFPropagateMetadataD3D11))
            clipped.append((int(b.has_clipped_path_info), [int(pm[i].clip_path_index) for i in range(b.path_count)]))

    scene.build(api.BuildOptions(), listener)
    assert kinds.count("PrepareClipTilesD3D11") == 1
    assert kinds.index("PrepareClipTilesD3D11") < kinds.index("DrawTilesD3D11")
    none = 0xFFFFFFFF
    assert clipped == [(1, [none, 0, 0])]


def test_porter_duff_identities_of_the_destructive_modes():
    """The twelve Porter-Duff operators partition source and destination: SrcIn + SrcOut = Copy, DestIn + DestOut =
    the destination, SrcOver = Copy + DestOut, DestAtop = SrcOut + DestIn (all on premultiplied colour, unclamped
    operands in [0, 1]); and nothing changes outside the drawn tiles."""
    rng = np.random.default_rng(11)
    dest = rng.random((8, 8, 4)).astype(np.float32)
    dest[..., :3] *= dest[..., 3:4]
    color = rng.random((8, 8, 4)).astype(np.float32)
    mask = rng.random((8, 8)).astype(np.float32)
    everywhere = np.ones((8, 8), bool)
    b = lambda mode: P.blend(dest, color, mask, mode, drawn=everywhere)
    assert np.allclose(b("src_in") + b("src_out"), b("copy"), atol=1e-6)
    assert np.allclose(b("dest_in") + b("dest_out"), dest, atol=1e-6)
    assert np.allclose(b("copy") + b("dest_out"), b("src_over"), atol=1e-6)
    assert np.allclose(b("src_out") + b("dest_in"), b("dest_atop"), atol=1e-6)
    assert not b("clear").any()
    drawn = np.zeros((8, 8), bool)
    drawn[:4] = True
    half = P.blend(dest, color, mask, "copy", drawn=drawn)
    assert np.array_equal(half[4:], dest[4:]) and np.array_equal(half[:4], b("copy")[:4])
    with pytest.raises(ValueError):
        P.blend(dest, color, mask, "copy")
