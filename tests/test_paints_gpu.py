"""GPU parity for SURVEY.md §8 f4: gradients, image patterns, pattern filters and blend modes.

Expected frames are chains of the two oracles: oracle/pf_oracle tiles and composites what lies under the paths being
tested and gives every tested path's mask (the path alone, opaque white on transparent black: its alpha channel);
oracle/paint_oracle.py evaluates the paint per pixel and blends it (numpy restatement of tile_fragment.inc.glsl +
gpu/blend.rs). Tolerance 1/255 for the frame underneath + 1/255 for the mask, as a byte difference <= 2, with the
mean error far below one level."""
import numpy as np
import pytest

from oracle import paint_oracle as P
from pathfinder_b200 import api
from pathfinder_b200.flat_scene import SceneBuilderPy
from tests import helpers as H

pytestmark = pytest.mark.gpu

SIZE = 256
WHITE = (1.0, 1.0, 1.0, 1.0)


def polygon(points, rgba=(255, 255, 255, 255), view=SIZE):
    b = SceneBuilderPy((0, 0, view, view))
    b.move_to(*points[0])
    for p in points[1:]:
        b.line_to(*p)
    b.close()
    b.end_path(rgba)
    return b.finish("polygon")


def blob(cx, cy, r, wobble=0.25, n=24, phase=0.0):
    """A wavy closed polygon: edges at every angle, sub-pixel vertices."""
    t = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    rad = r * (1.0 + wobble * np.sin(3.0 * t + phase))
    return [(float(cx + rad[i] * np.cos(t[i]) + 0.37), float(cy + rad[i] * np.sin(t[i]) + 0.21)) for i in range(n)]


def under_scene():
    b = SceneBuilderPy((0, 0, SIZE, SIZE))
    for pts, rgba in ((blob(90, 100, 70), (220, 60, 40, 255)), (blob(170, 150, 60, phase=1.0), (40, 90, 200, 160)),
                      ([(10.5, 200.25), (250.0, 180.0), (240.5, 250.75), (20.0, 240.0)], (30, 160, 70, 255))):
        b.move_to(*pts[0])
        for p in pts[1:]:
            b.line_to(*p)
        b.close()
        b.end_path(rgba)
    return b.finish("under")


def mask_of(points, area_lut, xf=None):
    """The path alone, opaque white on transparent black, through the oracle: alpha = mask alpha (float32)."""
    _, f = H.oracle_build(polygon(points), xf).render(area_lut, SIZE, SIZE, background=(0.0, 0.0, 0.0, 0.0), want_f32=True)
    return f[..., 3]


def drawn_of(points, xf=None):
    """Pixels of the tiles the path draws: the tile records of the path alone (empty tiles are never packed,
    builder.rs:1014-1016)."""
    drawn = np.zeros((SIZE, SIZE), bool)
    for t in H.oracle_build(polygon(points), xf).tiles:
        x, y = int(t["tile_x"]) * 16, int(t["tile_y"]) * 16
        drawn[max(y, 0):y + 16, max(x, 0):x + 16] = True
    return drawn


def run(under, paths, area_lut, xf=None, background=WHITE):
    """paths = [(points, push_paint(scene) -> paint id, color_fn(width, height) -> colour array or tuple, blend)].
    Returns (GPU frame RGBA8, expected frame float32)."""
    scene = api.Scene()
    scene.set_view_box((0.0, 0.0, float(SIZE), float(SIZE)))
    if under is not None:
        scene.push_flat(under)
    for points, push_paint, _, blend in paths:
        pts = np.asarray(points, np.float32)
        scene.push_draw_path(pts, np.zeros(len(pts), np.uint8), np.asarray([0, len(pts)], np.uint32), push_paint(scene),
                             blend_mode=api.BLEND_MODES[blend])
    r = api.CudaRenderer((SIZE, SIZE), background_color=background)
    t = None if xf is None else api.Transform2F(*xf)
    scene.build_and_render(r, api.BuildOptions(transform=t))
    img = r.read_pixels()
    r.close()
    if under is not None:
        _, dest = H.oracle_build(under, xf).render(area_lut, SIZE, SIZE, background=background, want_f32=True)
    else:
        dest = np.broadcast_to(np.asarray(background, np.float32), (SIZE, SIZE, 4)).copy()
    for points, _, color_fn, blend in paths:
        dest = P.blend(dest, color_fn(SIZE, SIZE), mask_of(points, area_lut, xf), blend,
                       drawn=drawn_of(points, xf) if blend in P.DESTRUCTIVE else None)
    return img, dest


def check(img, want, tol=2, mean_tol=0.35):
    want8 = np.clip(np.rint(np.clip(want, 0.0, 1.0) * 255.0), 0, 255).astype(np.int32)
    diff = np.abs(img.astype(np.int32) - want8)
    assert diff.max() <= tol, f"max RGBA diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}"
    assert diff.mean() <= mean_tol, f"mean RGBA diff {diff.mean():.3f}"


def solid(rgba):
    base = P.f16(np.asarray(rgba, np.float32) / np.float32(255.0))
    return (lambda scene: scene.push_paint(rgba)), (lambda w, h: base)


SMOOTH_MODES = ["src_over", "dest_over", "dest_out", "src_atop", "xor", "lighter", "darken", "lighten", "multiply", "screen",
                "hard_light", "overlay", "soft_light", "difference", "exclusion"]


@pytest.mark.parametrize("mode", SMOOTH_MODES)
def test_blend_modes(area_lut, mode):
    push_a, color_a = solid((250, 200, 30, 255))
    push_b, color_b = solid((60, 120, 240, 150))
    paths = [(blob(120, 110, 80, phase=2.0), push_a, color_a, mode), (blob(150, 160, 75, phase=0.5), push_b, color_b, mode)]
    # a transparent frame exercises the destination-alpha terms of the Porter-Duff modes
    background = (0.0, 0.0, 0.0, 0.0) if mode in ("dest_over", "src_atop", "xor", "dest_out") else WHITE
    img, want = run(under_scene(), paths, area_lut, background=background)
    check(img, want)
    # the tested paths changed the frame
    plain, _ = run(under_scene(), [], area_lut, background=background)
    assert (np.abs(img.astype(np.int32) - plain.astype(np.int32)).max(axis=2) > 8).mean() > 0.05


@pytest.mark.parametrize("mode", ["color_dodge", "color_burn", "hue", "saturation", "color", "luminosity"])
def test_blend_modes_with_steep_functions(area_lut, mode):
    """Dodge / burn divide by (1 - s), the HSL modes branch on the largest channel: a level of error underneath can
    become several on top, so these are held to the same bound on all but a sliver of the pixels."""
    push_a, color_a = solid((200, 150, 60, 255))
    push_b, color_b = solid((90, 140, 210, 200))
    paths = [(blob(120, 110, 80, phase=2.0), push_a, color_a, mode), (blob(150, 160, 75, phase=0.5), push_b, color_b, mode)]
    img, want = run(under_scene(), paths, area_lut)
    want8 = np.clip(np.rint(np.clip(want, 0.0, 1.0) * 255.0), 0, 255).astype(np.int32)
    diff = np.abs(img.astype(np.int32) - want8).max(axis=2)
    assert (diff > 2).mean() < 0.002, f"{(diff > 2).mean():.4f} of the pixels differ by more than 2 levels (max {diff.max()})"
    assert diff.mean() < 0.3


def gradient_paint(stops, line, row, radii=None, transform=None, repeat=False, render_transform=P.IDENTITY, other_ramps=()):
    """(push_paint, color_fn) for a gradient that is row `row` of the scene's gradient page."""
    def push(scene):
        return scene.push_gradient(stops, line, radii=radii, transform=transform, repeat=repeat)

    def color(w, h):
        ramps = list(other_ramps[:row]) + [P.gradient_ramp(stops)] + list(other_ramps[row:])
        page = P.gradient_page(ramps)
        flags = P.REPEAT_U if repeat else 0
        if radii is None:
            u, v = P.pixel_uv(P.linear_gradient_transform(line, row, render_transform), w, h)
            c = P.sample(page, u, v, flags)
        else:
            t, p0, p1 = P.radial_gradient_entry(line, radii, row, transform or P.IDENTITY, render_transform)
            u, v = P.pixel_uv(t, w, h)
            c = P.filter_radial_gradient(page, u, v, p0, p1, flags)
        return P.combine_src_in(c, (255, 255, 255, 255))
    return push, color


STOPS = [(0.0, (255, 40, 40, 255)), (0.35, (250, 240, 60, 255)), (0.7, (40, 200, 120, 128)), (1.0, (30, 60, 220, 255))]


@pytest.mark.parametrize("repeat", [False, True])
def test_linear_gradient(area_lut, repeat):
    push, color = gradient_paint(STOPS, ((60.0, 40.0), (150.0, 120.0)), 0, repeat=repeat)
    img, want = run(under_scene(), [(blob(128, 128, 100, wobble=0.15), push, color, "src_over")], area_lut)
    check(img, want)


def test_linear_gradient_under_a_build_transform(area_lut):
    xf = (0.8, 0.3, -0.2, 0.9, 30.0, 10.0)  # m11, m12, m21, m22, tx, ty as the scenes module hands transforms around
    t = api.Transform2F(*xf)
    render_transform = P.t_inverse((t.m11, t.m12, t.m21, t.m22, t.tx, t.ty))
    push, color = gradient_paint(STOPS, ((40.0, 60.0), (200.0, 180.0)), 0, render_transform=render_transform)
    img, want = run(under_scene(), [(blob(128, 128, 90, wobble=0.15), push, color, "src_over")], area_lut, xf=xf)
    check(img, want)


def test_two_gradients_and_a_blend_mode(area_lut):
    other = [(0.0, (255, 255, 255, 255)), (1.0, (0, 0, 0, 255))]
    push0, color0 = gradient_paint(STOPS, ((20.0, 20.0), (230.0, 60.0)), 0, other_ramps=[P.gradient_ramp(other)])
    push1, color1 = gradient_paint(other, ((128.0, 30.0), (128.0, 220.0)), 1, other_ramps=[P.gradient_ramp(STOPS)])
    paths = [(blob(110, 120, 90, wobble=0.1), push0, color0, "src_over"), (blob(150, 140, 80, phase=1.3), push1, color1, "multiply")]
    img, want = run(under_scene(), paths, area_lut)
    check(img, want)


@pytest.mark.parametrize("repeat", [False, True])
def test_radial_gradient(area_lut, repeat):
    push, color = gradient_paint(STOPS, ((120.0, 120.0), (140.0, 130.0)), 0, radii=(10.0, 90.0), repeat=repeat,
                                 transform=(1.0, 0.2, 0.0, 0.8, 5.0, 12.0))
    img, want = run(under_scene(), [(blob(128, 128, 105, wobble=0.1), push, color, "src_over")], area_lut)
    # t has a square root in it and the texture coordinate passes through f16: a band of pixels where adjacent ramp
    # texels differ may land one texel apart
    want8 = np.clip(np.rint(np.clip(want, 0.0, 1.0) * 255.0), 0, 255).astype(np.int32)
    diff = np.abs(img.astype(np.int32) - want8).max(axis=2)
    assert (diff > 2).mean() < 0.002 and diff.mean() < 0.3, ((diff > 2).mean(), diff.mean(), diff.max())


def image_paint(pixels, transform=None, repeat_x=False, repeat_y=False, smoothing=True, pattern_filter=None):
    def push(scene):
        return scene.push_image_pattern(pixels, transform=transform, repeat_x=repeat_x, repeat_y=repeat_y, smoothing=smoothing,
                                        pattern_filter=pattern_filter)

    def color(w, h):
        page = P.image_page(pixels, repeat_x, repeat_y)
        flags = (P.REPEAT_U if repeat_x else 0) | (P.REPEAT_V if repeat_y else 0) | (0 if smoothing else P.NEAREST)
        u, v = P.pixel_uv(P.image_pattern_transform((page.shape[1], page.shape[0]), transform or P.IDENTITY), w, h)
        if pattern_filter is None:
            c = P.sample(page, u, v, flags)
        elif "blur" in pattern_filter:
            c = P.filter_blur(page, u, v, pattern_filter["blur"], pattern_filter.get("direction", "x") == "y", flags)
        else:
            c = P.filter_color_matrix(page, u, v, pattern_filter["color_matrix"], flags)
        return P.combine_src_in(c, (255, 255, 255, 255))
    return push, color


def make_image(rng_seed=11):
    rng = np.random.default_rng(rng_seed)
    img = rng.integers(0, 256, (12, 16, 4), dtype=np.uint8)
    img[..., 3] = np.where(rng.random((12, 16)) < 0.3, rng.integers(60, 255, (12, 16)), 255)
    return img


@pytest.mark.parametrize("case", ["smooth", "nearest", "repeat", "repeat_nearest"])
def test_image_pattern(area_lut, case):
    pixels = make_image()
    kw = {"smooth": {}, "nearest": {"smoothing": False}, "repeat": {"repeat_x": True, "repeat_y": True},
          "repeat_nearest": {"repeat_x": True, "repeat_y": True, "smoothing": False}}[case]
    # 6.5x magnification, rotated a little: every pixel lands between texel centres
    push, color = image_paint(pixels, transform=(6.5, 1.0, -0.75, 6.0, 40.0, 50.0), **kw)
    img, want = run(under_scene(), [(blob(128, 128, 100, wobble=0.1), push, color, "src_over")], area_lut)
    if "nearest" in case:
        # a pixel within float rounding of a texel boundary may pick the neighbour: compare off the boundaries
        want8 = np.clip(np.rint(np.clip(want, 0.0, 1.0) * 255.0), 0, 255).astype(np.int32)
        diff = np.abs(img.astype(np.int32) - want8).max(axis=2)
        assert (diff > 2).mean() < 0.002, (diff > 2).mean()
    else:
        check(img, want)


@pytest.mark.parametrize("pattern_filter", [{"blur": 2.5, "direction": "x"}, {"blur": 1.2, "direction": "y"},
                                            {"color_matrix": [0.3, 0.3, 0.3, 0.0, 0.6, 0.6, 0.6, 0.0, 0.1, 0.1, 0.1, 0.0,
                                                              0.0, 0.0, 0.0, 1.0, 0.05, 0.0, 0.1, 0.0]}])
def test_pattern_filters(area_lut, pattern_filter):
    pixels = make_image(5)
    push, color = image_paint(pixels, transform=(8.0, 0.0, 0.0, 8.0, 30.0, 40.0), repeat_x=True, repeat_y=True,
                              pattern_filter=pattern_filter)
    img, want = run(under_scene(), [(blob(128, 128, 100, wobble=0.1), push, color, "src_over")], area_lut)
    check(img, want)


@pytest.mark.parametrize("mode", sorted(P.DESTRUCTIVE))
def test_destructive_blend_modes(area_lut, mode):
    """Clear, Copy, SrcIn, DestIn, SrcOut, DestAtop (effects.rs:222-235) change the pixels their mask leaves out too —
    on the tiles the path draws; the rest of the frame stays (builder.rs:1014-1016)."""
    push_a, color_a = solid((250, 200, 30, 255))
    push_b, color_b = solid((60, 120, 240, 150))
    paths = [(blob(120, 110, 80, phase=2.0), push_a, color_a, mode), (blob(150, 160, 75, phase=0.5), push_b, color_b, mode)]
    for background in ((0.0, 0.0, 0.0, 0.0), WHITE): # (destination alpha 0 outside the shapes underneath / 1 everywhere)
        img, want = run(under_scene(), paths, area_lut, background=background)
        check(img, want)
        plain, _ = run(under_scene(), [], area_lut, background=background)
        changed = np.abs(img.astype(np.int32) - plain.astype(np.int32)).max(axis=2) > 0
        assert changed.mean() > 0.02
        # tiles neither path draws keep what was there
        assert not changed[~(drawn_of(paths[0][0]) | drawn_of(paths[1][0]))].any()


def clipped_polygon(points, clip_points, view=SIZE):
    b = SceneBuilderPy((0, 0, view, view))
    b.move_to(*clip_points[0])
    for p in clip_points[1:]:
        b.line_to(*p)
    b.close()
    clip = b.end_clip_path()
    b.move_to(*points[0])
    for p in points[1:]:
        b.line_to(*p)
    b.close()
    b.end_path((255, 255, 255, 255), clip=clip)
    return b.finish("clipped polygon")


@pytest.mark.parametrize("mode", ["src_over", "multiply"])
def test_clipped_path_with_a_gradient(area_lut, mode):
    """f1 x f4: a clip path on a path with a textured paint (and a blend mode), in a scene with a display-list build."""
    points, clip_points = blob(128, 128, 100, wobble=0.15), blob(150, 120, 70, wobble=0.3, phase=1.0)
    under = under_scene()
    push, color = gradient_paint(STOPS, ((60.0, 40.0), (190.0, 200.0)), 0)
    scene = api.Scene()
    scene.set_view_box((0.0, 0.0, float(SIZE), float(SIZE)))
    scene.push_flat(under)
    cp = np.asarray(clip_points, np.float32)
    clip_id = scene.push_clip_path(cp, np.zeros(len(cp), np.uint8), np.asarray([0, len(cp)], np.uint32))
    pts = np.asarray(points, np.float32)
    scene.push_draw_path(pts, np.zeros(len(pts), np.uint8), np.asarray([0, len(pts)], np.uint32), push(scene),
                         blend_mode=api.BLEND_MODES[mode], clip_path_id=clip_id)
    r = api.CudaRenderer((SIZE, SIZE), background_color=WHITE)
    scene.build_and_render(r, api.BuildOptions())
    img = r.read_pixels()
    r.close()
    _, dest = H.oracle_build(under, None).render(area_lut, SIZE, SIZE, background=WHITE, want_f32=True)
    _, f = H.oracle_build(clipped_polygon(points, clip_points), None).render(area_lut, SIZE, SIZE, background=(0.0, 0.0, 0.0, 0.0),
                                                                            want_f32=True)
    mask = f[..., 3]
    assert 0.05 < (mask > 0.5).mean() < (mask_of(points, area_lut) > 0.5).mean() - 0.05  # the clip removes a good part
    check(img, P.blend(dest, color(SIZE, SIZE), mask, mode))
