"""CPU restatement of the text filter — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module. It
must never be imported from pathfinder_b200/.

Restates `filterText` and its helpers (shaders/tile_fragment.inc.glsl:91-166) in numpy float32 for the way the
reference's demo uses them (demo/common/src/lib.rs:804-834): the page's glyphs are drawn into a render target three
times as wide as the page (`BuildOptions::subpixel_aa_enabled` scales x by 3, renderer/src/scene.rs:260-262), and
one page-sized rectangle painted with that render target as a pattern runs the filter, so that page pixel (x, y)
samples the render target at texel centre 3x + 1.5.

PARITY UNPINNED: the reference holds no golden image or unit test for the filter. The sampler behaviour assumed here
is stated where it matters:
  * colour texture taps are `colorTexCoord + k / textureWidth`, k = -4..4 (:101-118): texel centres again, so a
    linear and a nearest sampler agree and tap k of pixel x is texel 3x + 1 + k; outside the texture the sampler
    clamps to the edge (patterns without the repeat flags, renderer/src/paint.rs pattern sampling flags);
  * the gamma table is sampled with bilinear filtering, clamp to edge, at (alpha, 1 - bgColor) (:122-124): texel
    coordinates u * 256 - 0.5 and v * 8 - 0.5.
Only the red channel of the render target is read (:93-95,147): the glyphs must be drawn in a colour whose red
channel is 1 (white) for it to be coverage.
"""
from __future__ import annotations

import numpy as np

_f = np.float32

# content/src/effects.rs:22-27
DEFRINGING_KERNEL_CORE_GRAPHICS = (0.033165660, 0.102074051, 0.221434336, 0.286651906)
DEFRINGING_KERNEL_FREETYPE = (0.0, 0.031372549, 0.301960784, 0.337254902)


def _tap(red: np.ndarray, k: int) -> np.ndarray:
    """filterTextSample1Tap (:93-95) for every page pixel: texel 3x + 1 + k of each row, clamped to the edge."""
    h, w3 = red.shape
    x = 3 * np.arange(w3 // 3) + 1 + k
    return red[:, np.clip(x, 0, w3 - 1)]


def _dot4(a, kernel):
    # GLSL dot(vec4, vec4): the order of the additions is implementation-defined; left to right here
    return ((a[0] * kernel[0] + a[1] * kernel[1]) + a[2] * kernel[2]) + a[3] * kernel[3]


def _dot3(a, kernel):
    return (a[0] * kernel[0] + a[1] * kernel[1]) + a[2] * kernel[2]


def _convolve7(alpha0, alpha1, kernel):
    """filterTextConvolve7Tap (:118-120): dot(alpha0, kernel) + dot(alpha1, kernel.zyx)."""
    return _dot4(alpha0, kernel) + _dot3(alpha1, (kernel[2], kernel[1], kernel[0]))


def sample_gamma_lut(gamma_lut: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """texture(gammaLUT, vec2(u, v)).r with a bilinear, clamp-to-edge sampler over the (8, 256) u8 table."""
    rows, cols = gamma_lut.shape
    table = gamma_lut.astype(_f) / _f(255.0)
    x = np.asarray(u, _f) * _f(cols) - _f(0.5)
    y = np.asarray(v, _f) * _f(rows) - _f(0.5)
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0).astype(_f), (y - y0).astype(_f)
    x0i, y0i = x0.astype(np.int64), y0.astype(np.int64)
    xa, xb = np.clip(x0i, 0, cols - 1), np.clip(x0i + 1, 0, cols - 1)
    ya, yb = np.clip(y0i, 0, rows - 1), np.clip(y0i + 1, 0, rows - 1)
    top = table[ya, xa] * (_f(1.0) - fx) + table[ya, xb] * fx
    bottom = table[yb, xa] * (_f(1.0) - fx) + table[yb, xb] * fx
    return (top * (_f(1.0) - fy) + bottom * fy).astype(_f)


def filter_text(red: np.ndarray, fg_color, bg_color, defringing_kernel=DEFRINGING_KERNEL_CORE_GRAPHICS,
                gamma_lut: np.ndarray | None = None) -> np.ndarray:
    """filterText (:134-166) over a whole page.

    red: (H, 3W) float32, the red channel of the render target as the sampler returns it (0..1; an RGBA8 target
    holds multiples of 1/255). With defringing_kernel = None the render target is page-sized, (H, W), and sampled
    once per pixel (kernel.w == 0, :146-147). fg_color / bg_color: RGB in 0..1 (filterParams2 / filterParams1).
    gamma_lut: the (8, 256) u8 table to enable gamma correction (filterParams2.a != 0), else None.
    Returns (H, W, 4) float32 = vec4(mix(bgColor, fgColor, alpha), 1)."""
    red = np.asarray(red, _f)
    fg = np.asarray(fg_color, _f)[:3]
    bg = np.asarray(bg_color, _f)[:3]
    kernel = None if defringing_kernel is None else tuple(_f(k) for k in defringing_kernel)
    if kernel is None or kernel[3] == 0.0:
        if kernel is not None:  # a 3x-wide target with a zero kernel: centre taps only
            red = _tap(red, 0)
        alpha = np.stack([red, red, red], axis=-1)
    else:
        wide = kernel[0] > 0.0  # filterTextSample9Tap (:99-116): the outermost taps only for a 4-weight kernel
        zero = np.zeros_like(_tap(red, 0))
        left = [_tap(red, -4) if wide else zero, _tap(red, -3), _tap(red, -2), _tap(red, -1)]
        center = _tap(red, 0)
        right = [_tap(red, 1), _tap(red, 2), _tap(red, 3), _tap(red, 4) if wide else zero]
        r = _convolve7(left, (center, right[0], right[1]), kernel)
        g = _convolve7((left[1], left[2], left[3], center), (right[0], right[1], right[2]), kernel)
        b = _convolve7((left[2], left[3], center, right[0]), (right[1], right[2], right[3]), kernel)
        alpha = np.stack([r, g, b], axis=-1).astype(_f)
    if gamma_lut is not None:
        # filterTextGammaCorrect (:122-132): per channel, texture(gammaLUT, vec2(alpha, 1 - bgColor)).r
        alpha = np.stack([sample_gamma_lut(gamma_lut, alpha[..., c], np.broadcast_to(_f(1.0) - bg[c], alpha[..., c].shape))
                          for c in range(3)], axis=-1)
    rgb = bg + (fg - bg) * alpha  # mix(bgColor, fgColor, alpha) = x * (1 - a) + y * a, written as the usual lowering
    return np.concatenate([rgb.astype(_f), np.ones(rgb.shape[:-1] + (1,), _f)], axis=-1)
