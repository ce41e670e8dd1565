"""GPU parity for SURVEY.md §8 f3, device side: render targets (PushRenderTarget / PopRenderTarget), pattern paints
over them, and the text filter (9-tap defringing + gamma LUT) — BASELINE.json configs[2] the way the reference's demo
wraps a text page (demo/common/src/lib.rs:804-834). The expected frame is the CPU chain of
tests/test_text_filter_oracle.py: the oracle tiles and composites the 3x-wide target, oracle/text_filter.py filters it."""
import numpy as np
import pytest

from oracle import text_filter as T
from pathfinder_b200 import api, gamma_lut, scenes
from tests import helpers as H

pytestmark = pytest.mark.gpu

KERNEL = T.DEFRINGING_KERNEL_CORE_GRAPHICS


def f16(values):
    """The reference ships paint parameters through an RGBA16F metadata texture (gpu/renderer.rs:712-763)."""
    return tuple(float(v) for v in np.asarray(values, np.float32).astype(np.float16).astype(np.float32))


def oracle_target(wide, size, area_lut, dilation):
    """The 3x-wide render target as the oracle tiles and composites it (RGBA8, transparent background)."""
    return H.oracle_build(wide, None, dilation=dilation).render(area_lut, 3 * size, size)


def filtered(target, fg, bg, kernel, gamma):
    """oracle/text_filter.py over a render target's red channel -> the page, RGBA8."""
    red = target[:, :, 0].astype(np.float32) / np.float32(255.0)
    page = T.filter_text(red, f16(fg), f16(bg), defringing_kernel=None if kernel is None else f16(kernel),
                         gamma_lut=gamma_lut.generate() if gamma else None)
    return np.clip(np.rint(page * 255.0), 0, 255).astype(np.uint8)


def check_page(scene, wide, size, area_lut, dilation, fg, bg, kernel, gamma, background=(0.5, 0.5, 0.5, 1.0)):
    """Two links, each within 1/255: the render target against the oracle's, and the page against the oracle's
    filter applied to the target the GPU actually produced. (The gamma table is steep — up to 3 output levels per
    input level — so comparing the page with a filter of the ORACLE's target would compound the two tolerances.)"""
    r = api.CudaRenderer((size, size), background_color=background)
    scene.build_and_render(r, api.BuildOptions(dilation=dilation))
    img = r.read_pixels()
    target = r.read_texture_page(0)
    want_target = oracle_target(wide, size, area_lut, dilation)
    assert target.shape == want_target.shape
    assert np.abs(target.astype(np.int32) - want_target.astype(np.int32)).max() <= 1, "render target differs from the oracle's"
    want = filtered(target, fg, bg, kernel, gamma)
    diff = np.abs(img.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1, f"max RGBA diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}"
    assert (img[..., 3] == 255).all()
    return r, img


@pytest.mark.parametrize("gamma", [False, True])
@pytest.mark.parametrize("colors", [((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), ((0.9, 0.85, 0.2), (0.1, 0.15, 0.3))])
def test_text_page_with_subpixel_aa(area_lut, gamma, colors):
    fg, bg = colors
    n, size = 600, 512
    wide = scenes.text_page_subpixel(n, size, layout="lines")
    dilation = (0.0121 * 16 * 3, 0.0121 * 1.25 * 16)  # STEM_DARKENING_FACTORS * font size, x in subpixels
    scene = scenes.subpixel_scene(wide, size, fg=fg, bg=bg, kernel=KERNEL, gamma=gamma)
    r, img = check_page(scene, wide, size, area_lut, dilation, fg, bg, KERNEL, gamma)
    # the page is not blank, and a second frame (render target redrawn from transparent black) is identical
    assert (np.abs(img[..., :3].astype(np.int32) - np.rint(np.asarray(bg) * 255)).max(axis=2) > 64).mean() > 0.01
    scene.build_and_render(r, api.BuildOptions(dilation=dilation))
    assert np.array_equal(r.read_pixels(), img)
    r.close()


def test_text_page_full_size_config3(area_lut):
    """BASELINE.json configs[2] at its full size: 10,000 glyphs at 12-16 px on 2048 x 2048, subpixel AA, stem
    darkening, gamma LUT."""
    n, size = 10000, 2048
    wide = scenes.text_page_subpixel(n, size)
    dilation = (0.0121 * 16 * 3, 0.0121 * 1.25 * 16)
    scene = scenes.subpixel_scene(wide, size)
    r, _ = check_page(scene, wide, size, area_lut, dilation, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), KERNEL, True,
                      background=(1.0, 1.0, 1.0, 1.0))
    r.close()


def test_unfiltered_render_target_pattern(area_lut):
    """A render target composited back through a plain pattern paint (no filter), over a background path: the
    target's pixels (premultiplied RGBA8) are the paint's colour, blended like any other colour
    (filterNone + combineColor0 SrcIn + premultiply, shaders/tile_fragment.inc.glsl:361-363,81-89,611)."""
    size = 256
    inner = scenes.random_paths(40, size, 5, r_min=8.0, r_max=60.0)
    opaque = inner.paint_colors.copy()
    opaque[:, 3] = 255
    inner.paint_colors = opaque
    want_target = H.oracle_build(inner, None).render(area_lut, size, size)  # what the render target should hold

    # what lies under the pattern: a blue shape on white, rendered by the oracle
    from pathfinder_b200.flat_scene import SceneBuilderPy
    b = SceneBuilderPy((0, 0, size, size))
    b.move_to(20.5, 10.25), b.line_to(200.0, 30.0), b.line_to(170.5, 190.75), b.line_to(5.0, 150.0)
    b.close()
    b.end_path((30, 60, 200, 255))
    under = b.finish("under")
    _, dest = H.oracle_build(under, None).render(area_lut, size, size, background=(1.0, 1.0, 1.0, 1.0), want_f32=True)

    scene = api.Scene()
    scene.set_view_box((0.0, 0.0, float(size), float(size)))
    scene.push_flat(under)
    rt = scene.push_render_target(size, size)
    scene.push_flat(inner)
    scene.pop_render_target()
    paint = scene.push_render_target_pattern(rt)
    # the pattern's rectangle overhangs the frame: every tile inside is a solid tile of that path (an edge ON a tile
    # boundary would give the last pixel column of each tile 255/256 coverage — fills are clamped to 4095/256)
    m = 24.0
    rect = np.asarray([[-m, -m], [size + m, -m], [size + m, size + m], [-m, size + m]], np.float32)
    scene.push_draw_path(rect, np.zeros(4, np.uint8), np.asarray([0, 4], np.uint32), paint)
    r = api.CudaRenderer((size, size), background_color=(1.0, 1.0, 1.0, 1.0))
    scene.build_and_render(r, api.BuildOptions())
    img = r.read_pixels().astype(np.float32) / 255.0
    # link 1: the render target itself against the oracle; link 2: the composite of the target the GPU produced
    target = r.read_texture_page(0)
    assert np.abs(target.astype(np.int32) - want_target.astype(np.int32)).max() <= 1
    t = target.astype(np.float32) / 255.0
    a = t[..., 3:4]
    want = dest * (1.0 - a) + np.concatenate([t[..., :3] * a, a], axis=2)  # colour = texel, alpha = texel alpha, premultiplied again
    # 1/255 for the frame under the pattern (the oracle's, not the GPU's) + half a level of RGBA8 rounding
    assert np.abs(img - want).max() <= 1.5 / 255.0 + 1e-6
    assert (a > 0).mean() > 0.05 and (a < 1).mean() > 0.3  # the pattern is neither empty nor everywhere
    r.close()


def test_render_target_protocol_errors():
    from pathfinder_b200 import _lib as L
    r = api.CudaRenderer((64, 64))
    r.begin_scene()
    cmd = L.PFRenderCommand()
    cmd.kind = L.PF_RENDER_COMMAND_POP_RENDER_TARGET
    with pytest.raises(L.PathfinderCudaError) as e:
        r.render_command(cmd)
    assert e.value.status == L.PF_CUDA_ERROR_PROTOCOL
    cmd.kind = L.PF_RENDER_COMMAND_PUSH_RENDER_TARGET
    cmd.u.push_render_target.render_target_id = 3
    with pytest.raises(L.PathfinderCudaError) as e:
        r.render_command(cmd)
    assert e.value.status == L.PF_CUDA_ERROR_PROTOCOL
    cmd.kind = L.PF_RENDER_COMMAND_DECLARE_RENDER_TARGET
    cmd.u.declare_render_target.render_target_id = 0
    cmd.u.declare_render_target.location.page = 7  # never allocated
    with pytest.raises(L.PathfinderCudaError) as e:
        r.render_command(cmd)
    assert e.value.status == L.PF_CUDA_ERROR_INVALID_ARGUMENT
    r.end_scene()
    r.close()
