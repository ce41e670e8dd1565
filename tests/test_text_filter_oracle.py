"""oracle/text_filter.py: the numpy restatement of the reference's text filter (parity unpinned,
see its header). Properties of the filter itself, and the text page run through the whole CPU chain: stem-darkened
outlines tiled three times as wide as the page, then defringed and gamma-corrected back to page pixels."""
import numpy as np

from oracle import text_filter as T
from pathfinder_b200 import gamma_lut, scenes
from tests import helpers as H

BLACK, WHITE = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)


def test_kernels_are_normalised():
    total = lambda k: 2 * (k[0] + k[1] + k[2]) + k[3]
    assert abs(total(T.DEFRINGING_KERNEL_CORE_GRAPHICS) - 1.0) < 1e-6
    # FreeType's weights are (0, 8, 77, 86) / 255: they add up to 256 / 255 (content/src/effects.rs:26-27)
    assert abs(total(T.DEFRINGING_KERNEL_FREETYPE) - 256.0 / 255.0) < 1e-6


def test_flat_coverage_maps_to_the_two_colours():
    fg, bg = (0.1, 0.2, 0.3), (0.9, 0.8, 0.7)
    full = T.filter_text(np.ones((4, 30), np.float32), fg, bg)
    none = T.filter_text(np.zeros((4, 30), np.float32), fg, bg)
    assert full.shape == (4, 10, 4) and (full[..., 3] == 1).all()
    assert np.allclose(full[..., :3], fg, atol=1e-6) and np.allclose(none[..., :3], bg, atol=1e-6)
    # and through the gamma table too: its first and last columns are fixed points
    lut = gamma_lut.generate()
    assert np.allclose(T.filter_text(np.ones((4, 30), np.float32), fg, bg, gamma_lut=lut)[..., :3], fg, atol=2e-3)
    assert np.allclose(T.filter_text(np.zeros((4, 30), np.float32), fg, bg, gamma_lut=lut)[..., :3], bg, atol=2e-3)


def test_vertical_edge_gives_ordered_colour_fringes():
    """Ink from subpixel 15 on (page pixel 5 on): the red subpixel of a pixel is the leftmost, so near the edge
    alpha.r <= alpha.g <= alpha.b, and each channel is a step response of the 7-tap kernel."""
    red = np.zeros((1, 30), np.float32)
    red[:, 15:] = 1.0
    out = T.filter_text(red, BLACK, WHITE)  # black on white: channel value = 1 - alpha
    alpha = 1.0 - out[0, :, :3]
    assert (np.diff(alpha, axis=0) >= -1e-7).all()
    assert (alpha[:, 0] <= alpha[:, 1] + 1e-7).all() and (alpha[:, 1] <= alpha[:, 2] + 1e-7).all()
    assert np.allclose(alpha[2], 0) and np.allclose(alpha[8], 1, atol=1e-6)
    k = T.DEFRINGING_KERNEL_CORE_GRAPHICS
    # pixel 4's blue channel is centred on subpixel 14: taps 11..17, three of them (15, 16, 17) inked
    assert abs(alpha[4, 2] - (k[2] + k[1] + k[0])) < 1e-6
    # pixel 5's red channel is centred on subpixel 15: its centre tap and everything to the right
    assert abs(alpha[5, 0] - (k[3] + k[2] + k[1] + k[0])) < 1e-6


def test_freetype_kernel_skips_the_outermost_taps():
    red = np.zeros((1, 30), np.float32)
    red[0, 3 * 5 + 1 + 4] = 1.0  # only the +4 tap of pixel 5
    assert T.filter_text(red, BLACK, WHITE, T.DEFRINGING_KERNEL_FREETYPE)[0, 5, 2] == 1.0   # kernel.x == 0: not sampled
    assert T.filter_text(red, BLACK, WHITE, T.DEFRINGING_KERNEL_CORE_GRAPHICS)[0, 5, 2] < 1.0


def test_no_kernel_is_a_plain_mix():
    cov = np.linspace(0, 1, 12, dtype=np.float32).reshape(2, 6)
    out = T.filter_text(cov, (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), defringing_kernel=None)
    assert out.shape == (2, 6, 4)
    assert np.allclose(out[..., 0], cov) and np.allclose(out[..., 2], 1 - cov) and (out[..., 1] == 0).all()


def test_gamma_correction_thins_dark_text_and_thickens_light_text():
    lut = gamma_lut.generate()
    cov = np.full((1, 30), 0.5, np.float32)
    dark = T.filter_text(cov, BLACK, WHITE, gamma_lut=lut)[0, 5, :3]
    light = T.filter_text(cov, WHITE, BLACK, gamma_lut=lut)[0, 5, :3]
    assert (dark > 0.5).all()    # less than half of the black goes down on white
    assert (light > 0.5).all()   # more than half of the white goes down on black
    # bilinear sampling: alpha = 0.5 falls between columns 127 and 128 of the row for bg = 1 (row 0, clamped)
    want = (float(lut[0, 127]) + float(lut[0, 128])) / 2 / 255
    assert abs((1.0 - dark[0]) - want) < 1e-6


def test_text_page_through_the_cpu_chain(area_lut):
    """Config 3 on the CPU: white glyphs, x scaled by 3 into a 3W x H target, stem darkening for 16 px, defringed
    without gamma correction: the filter is normalised, so the page's total ink is a third of the target's."""
    n, size = 200, 256
    flat = scenes.text_page(n, size, layout="lines")
    wide = flat.with_view_box((0.0, 0.0, 3.0 * size, float(size)))
    wide.paint_colors = np.asarray([[255, 255, 255, 255]], np.uint8)
    dilation = (0.0121 * 16 * 3, 0.0121 * 1.25 * 16)
    built = H.oracle_build(wide, (3.0, 0.0, 0.0, 1.0, 0.0, 0.0), dilation=dilation)
    target = built.render(area_lut, 3 * size, size)           # RGBA8, transparent background
    red = target[:, :, 0].astype(np.float32) / np.float32(255.0)
    assert red.max() == 1.0 and red.min() == 0.0
    page = T.filter_text(red, BLACK, WHITE)
    assert page.shape == (size, size, 4)
    ink = (1.0 - page[..., :3]).sum(axis=(0, 1))
    assert np.allclose(ink, red.sum() / 3.0, rtol=2e-3)
    corrected = T.filter_text(red, BLACK, WHITE, gamma_lut=gamma_lut.generate())
    assert (1.0 - corrected[..., :3]).sum() < ink.sum()        # black on white gets lighter
