"""CPU restatement of the reference's paint / blend surface — TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module; the product (pathfinder_b200/) never does. PARITY UNPINNED: the reference holds
one known-answer test near this code (Gradient::sample, content/src/gradient.rs:291-303, restated in
tests/test_paint_oracle.py); everything else here is a reading of the Rust / GLSL sources, in numpy float32:

  Gradient::sample                       content/src/gradient.rs:188-211
  GradientTileBuilder (the 256-texel ramp)  renderer/src/paint.rs:813-873
  calculate_texture_transforms           renderer/src/paint.rs:597-639
  PaintMetadata::filter (uv_origin)      renderer/src/paint.rs:781-800
  compute_filter_params                  renderer/src/gpu/renderer.rs:967-1049 (through RGBA16F, :712-763)
  computeTileVaryings / filterColor / combineColor0 / composite / calculateColor
                                         shaders/tile_fragment.inc.glsl:81-89,274-412,414-535,560-614
  blend states of the Porter-Duff modes  renderer/src/gpu/blend.rs:43-163

Semantics are D3D9's (the D3D11 tile shader binds a placeholder as the destination texture, tile.cs.glsl:140-142,
and the palette never sets a blend mode, paint.rs:576): a path's pixels are calculateColor() of the current
destination pixel, then the blend state of its mode; for the modes the shader evaluates itself blending is off and
the result (alpha forced to 1) replaces the pixel. Sampler: LINEAR, CLAMP_TO_EDGE unless REPEAT / NEAREST flags.
"""
from __future__ import annotations

import numpy as np

F = np.float32

BLEND_MODES = ["clear", "copy", "src_in", "src_out", "src_over", "src_atop", "dest_in", "dest_out", "dest_over",
               "dest_atop", "xor", "lighter", "darken", "lighten", "multiply", "screen", "hard_light", "overlay",
               "color_dodge", "color_burn", "soft_light", "difference", "exclusion", "hue", "saturation", "color",
               "luminosity"]
DESTRUCTIVE = {"clear", "copy", "src_in", "dest_in", "src_out", "dest_atop"}  # BlendMode::is_destructive, effects.rs:222-235

REPEAT_U, REPEAT_V, NEAREST = 0x1, 0x2, 0x4 | 0x8


def f16(x):
    """Paint parameters travel through an RGBA16F metadata texture (gpu/renderer.rs:712-763)."""
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# Gradients
# ---------------------------------------------------------------------------------------------------------------

def sort_stops(stops):
    """Gradient::add_color_stop (gradient.rs:141-150): inserted after every stop whose offset is <= the new one."""
    out = []
    for offset, color in stops:
        index = 0
        while index < len(out) and out[index][0] <= offset:
            index += 1
        out.insert(index, (offset, color))
    return out


def gradient_sample(stops, t):
    """Gradient::sample (gradient.rs:188-211). stops = [(offset, (r, g, b, a))], sorted."""
    if not stops:
        return (0, 0, 0, 0)
    t = F(min(max(F(t), F(0.0)), F(1.0)))
    last = len(stops) - 1
    lo, hi = 0, len(stops)  # binary_search_by with the comparator that never returns Equal: a partition point
    while lo < hi:
        mid = lo + (hi - lo) // 2
        if F(stops[mid][0]) < t or F(stops[mid][0]) == F(0.0):
            lo = mid + 1
        else:
            hi = mid
    upper = min(lo, last)
    lower = upper - 1 if upper > 0 else upper
    (o0, c0), (o1, c1) = stops[lower], stops[upper]
    denom = F(F(o1) - F(o0))
    if denom == 0:
        return tuple(int(v) for v in c0)
    ratio = F(min(F(F(t - F(o0)) / denom), F(1.0)))
    a = np.asarray(c0, np.float32) * F(1.0 / 255.0)
    b = np.asarray(c1, np.float32) * F(1.0 / 255.0)
    v = (a + (b - a) * ratio) * F(255.0)
    return tuple(int(x) for x in np.rint(v).astype(np.int64))  # to_i32x4 = cvtps: round to nearest even


def gradient_ramp(stops):
    """GradientTileBuilder::allocate (paint.rs:859-863): 256 texels sampled at t = (x + 0.5) / 256."""
    return np.asarray([gradient_sample(stops, F((F(x) + F(0.5)) / F(256.0))) for x in range(256)], np.uint8)


def gradient_page(ramps):
    """The 256 x 256 tile: one gradient per row, the rest ColorU::black() (paint.rs:836-840)."""
    page = np.zeros((256, 256, 4), np.uint8)
    page[..., 3] = 255
    for row, ramp in enumerate(ramps):
        page[row] = ramp
    return page


# ---------------------------------------------------------------------------------------------------------------
# Texture transforms. A transform is (m11, m12, m21, m22, tx, ty): x' = m11 x + m12 y + tx, y' = m21 x + m22 y + ty.
# ---------------------------------------------------------------------------------------------------------------

IDENTITY = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)


def t_mul(a, b):
    """a * b: apply b first."""
    a = [F(v) for v in a]
    b = [F(v) for v in b]
    return (a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3], a[2] * b[0] + a[3] * b[2], a[2] * b[1] + a[3] * b[3],
            a[0] * b[4] + a[1] * b[5] + a[4], a[2] * b[4] + a[3] * b[5] + a[5])


def t_inverse(t):
    m11, m12, m21, m22, tx, ty = [F(v) for v in t]
    det = m11 * m22 - m12 * m21
    i11, i12, i21, i22 = m22 / det, -m12 / det, -m21 / det, m11 / det
    return (i11, i12, i21, i22, -(i11 * tx + i12 * ty), -(i21 * tx + i22 * ty))


def linear_gradient_transform(line, row, render_transform=IDENTITY):
    """paint.rs:611-618: the gradient line projected onto (0..1, v0), v0 = the centre of the gradient's row."""
    (x0, y0), (x1, y1) = [(F(p[0]), F(p[1])) for p in line]
    dx, dy = x1 - x0, y1 - y0
    len2 = dx * dx + dy * dy
    m0x, m0y = dx / len2, dy / len2
    v0 = (F(row) + F(0.5)) * F(1.0 / 256.0)
    return t_mul((m0x, m0y, F(0), F(0), m0x * -x0 + m0y * -y0, v0), render_transform)


def radial_gradient_entry(line, radii, row, transform=IDENTITY, render_transform=IDENTITY):
    """paint.rs:619-622 + 788-793 + gpu/renderer.rs:977-986: (texture transform, p0, p1)."""
    (x0, y0), (x1, y1) = line
    p0 = (x0, y0, F(x1) - F(x0), F(y1) - F(y0))
    p1 = (radii[0], radii[1], 0.0, (F(row) + F(0.5)) * F(1.0 / 256.0))
    return t_mul(t_inverse(transform), render_transform), p0, p1


def image_pattern_transform(page_size, pattern_transform=IDENTITY, render_transform=IDENTITY):
    """paint.rs:623-631: from_scale(texture_scale).translate(rect origin uv = 0) * pattern.transform().inverse().
    The page is the image plus a texel of border on each side it does not repeat on; the border is not added back."""
    sx, sy = F(1.0) / F(page_size[0]), F(1.0) / F(page_size[1])
    inv = t_inverse(pattern_transform)
    return t_mul((sx * inv[0], sx * inv[1], sy * inv[2], sy * inv[3], sx * inv[4], sy * inv[5]), render_transform)


def image_page(pixels, repeat_x, repeat_y):
    """The image's own page: a transparent texel of border on every side it does not repeat on (paint.rs:501-530)."""
    px = np.asarray(pixels, np.uint8)
    bx, by = (0 if repeat_x else 1), (0 if repeat_y else 1)
    page = np.zeros((px.shape[0] + 2 * by, px.shape[1] + 2 * bx, 4), np.uint8)
    page[by:by + px.shape[0], bx:bx + px.shape[1]] = px
    return page


# ---------------------------------------------------------------------------------------------------------------
# Sampling and filters. Textures are (h, w, 4) uint8, rows top-down; uv arrays are float32.
# ---------------------------------------------------------------------------------------------------------------

def _wrap(i, n, repeat):
    return np.mod(i, n) if repeat else np.clip(i, 0, n - 1)


def sample(texture, u, v, flags=0, bottom_up=False):
    """texture(colorTexture, uv) -> (..., 4) float32 in [0, 1]."""
    tex = np.asarray(texture, np.uint8)
    h, w = tex.shape[:2]
    texf = tex.astype(np.float32) * F(1.0 / 255.0)
    x = np.asarray(u, np.float32) * F(w) - F(0.5)
    y = (F(h) - F(0.5) - np.asarray(v, np.float32) * F(h)) if bottom_up else (np.asarray(v, np.float32) * F(h) - F(0.5))
    ru, rv = bool(flags & REPEAT_U), bool(flags & REPEAT_V)
    if flags & NEAREST:
        xi = _wrap(np.floor(x + F(0.5)).astype(np.int64), w, ru)
        yi = _wrap(np.floor(y + F(0.5)).astype(np.int64), h, rv)
        return texf[yi, xi]
    fx, fy = np.floor(x), np.floor(y)
    ax, ay = (x - fx)[..., None], (y - fy)[..., None]
    x0, x1 = _wrap(fx.astype(np.int64), w, ru), _wrap(fx.astype(np.int64) + 1, w, ru)
    y0, y1 = _wrap(fy.astype(np.int64), h, rv), _wrap(fy.astype(np.int64) + 1, h, rv)
    top = texf[y0, x0] + (texf[y0, x1] - texf[y0, x0]) * ax
    bottom = texf[y1, x0] + (texf[y1, x1] - texf[y1, x0]) * ax
    return (top + (bottom - top) * ay).astype(np.float32)


def filter_radial_gradient(texture, u, v, p0, p1, flags=0):
    """filterRadialGradient (tile_fragment.inc.glsl:274-300)."""
    p0, p1 = f16(p0), f16(p1)
    dpx, dpy = u - p0[0], v - p0[1]
    dcx, dcy = p0[2], p0[3]
    dr = p1[1] - p1[0]
    a = (dcx * dcx + dcy * dcy) - dr * dr
    b = (dpx * dcx + dpy * dcy) + p1[0] * dr
    c = (dpx * dpx + dpy * dpy) - p1[0] * p1[0]
    discrim = b * b - a * c
    with np.errstate(invalid="ignore", divide="ignore"):
        root = np.sqrt(discrim)
        t0, t1 = (root + b) / a, (-root + b) / a
    lo, hi = np.minimum(t0, t1), np.maximum(t0, t1)
    t = np.where(lo >= 0, lo, hi).astype(np.float32)
    color = sample(texture, np.nan_to_num(p1[2] + t, nan=0.0, posinf=1e9, neginf=-1e9).astype(np.float32),
                   np.full_like(t, p1[3]), flags)
    return np.where((discrim != 0)[..., None], color, F(0.0)).astype(np.float32)


def filter_blur(texture, u, v, sigma, vertical, flags=0, bottom_up=False):
    """filterBlur (tile_fragment.inc.glsl:302-339) with compute_filter_params' coefficients (gpu/renderer.rs:988-1009)."""
    h, w = np.asarray(texture).shape[:2]
    sigma_inv = F(1.0) / F(sigma)
    gy_raw = F(np.exp(F(-0.5) * sigma_inv * sigma_inv))
    gx, gy, gz = f16(F(0.3989422804014327) * sigma_inv), f16(gy_raw), f16(gy_raw * gy_raw)  # SQRT_2_PI_INV / sigma, ...
    support = int(f16(np.ceil(F(1.5) * F(sigma)) * F(2.0)))
    ox, oy = (F(0.0), F(1.0) / F(h)) if vertical else (F(1.0) / F(w), F(0.0))
    gx, gy, gz = F(gx), F(gy), F(gz)
    total = gx
    color = sample(texture, u, v, flags, bottom_up) * gx
    gx, gy = gx * gy, gy * gz
    for i in range(1, support + 1, 2):
        partial = gx
        gx, gy = gx * gy, gy * gz
        partial = partial + gx
        off = F(i) + gx / partial
        color = color + (sample(texture, u - ox * off, v - oy * off, flags, bottom_up) +
                         sample(texture, u + ox * off, v + oy * off, flags, bottom_up)) * partial
        total = total + F(2.0) * partial
        gx, gy = gx * gy, gy * gz
    return (color / total).astype(np.float32)


def filter_color_matrix(texture, u, v, columns, flags=0, bottom_up=False):
    """filterColorMatrix (:341-351): mat4(p0..p3) * colour + p4 — p0..p3 are the matrix's columns."""
    m = f16(columns).reshape(5, 4)
    c = sample(texture, u, v, flags, bottom_up)
    out = (c[..., 0:1] * m[0] + c[..., 1:2] * m[1] + c[..., 2:3] * m[2] + c[..., 3:4] * m[3]) + m[4]
    return out.astype(np.float32)


def pixel_uv(transform, width, height):
    """computeTileVaryings: the texture coordinate of every pixel centre, transform rounded through f16."""
    m11, m12, m21, m22, tx, ty = f16(transform)
    ys, xs = np.mgrid[0:height, 0:width]
    fx, fy = xs.astype(np.float32) + F(0.5), ys.astype(np.float32) + F(0.5)
    return (m11 * fx + m12 * fy + tx).astype(np.float32), (m21 * fx + m22 * fy + ty).astype(np.float32)


def combine_src_in(color0, base_rgba8):
    """combineColor0, SrcIn: (src.rgb, src.a * dest.a) with dest = the base colour through f16."""
    base = f16(np.asarray(base_rgba8, np.float32) * F(1.0 / 255.0))
    out = color0.copy()
    out[..., 3] = out[..., 3] * base[3]
    return out


# ---------------------------------------------------------------------------------------------------------------
# Blend modes
# ---------------------------------------------------------------------------------------------------------------

def _divide(num, denom):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(denom != 0, num / np.where(denom != 0, denom, 1), F(0.0)).astype(np.float32)


def _color_dodge(d, s):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(d == 0, F(0.0), np.where(s == 1, F(1.0), d / np.where(s == 1, F(1.0), F(1.0) - s))).astype(np.float32)


def _screen(d, s):
    return d + s - d * s


def _hard_light(d, s):
    return np.where(s <= F(0.5), d * F(2.0) * s, _screen(d, F(2.0) * s - F(1.0)))


def _soft_light(d, s):
    darkened = np.where(d <= F(0.25), ((F(16.0) * d - F(12.0)) * d + F(4.0)) * d, np.sqrt(np.maximum(d, 0)))
    factor = np.where(s <= F(0.5), d * (F(1.0) - d), darkened - d)
    return d + (s * F(2.0) - F(1.0)) * factor


def _rgb_to_hsl(rgb):
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    v = np.maximum(np.maximum(r, g), b)
    x_min = np.minimum(np.minimum(r, g), b)
    c = v - x_min
    l = x_min + (v - x_min) * F(0.5)
    t0 = np.where(r == v, F(0.0), np.where(g == v, F(2.0), F(4.0)))
    t1 = np.where(r == v, g, np.where(g == v, b, r))
    t2 = np.where(r == v, b, np.where(g == v, r, g))
    h = F(np.pi / 3.0) * _divide(t0 * c + t1 - t2, c)
    return np.stack([h, _divide(c, v), l], axis=-1).astype(np.float32)


def _hsl_to_rgb(hsl):
    h, s, l = hsl[..., 0], hsl[..., 1], hsl[..., 2]
    a = s * np.minimum(l, F(1.0) - l)
    out = []
    for n in (0.0, 8.0, 4.0):
        k = np.mod(F(n) + h * F(6.0 / np.pi), F(12.0))
        out.append(l - np.clip(np.minimum(k - F(3.0), F(9.0) - k), F(-1.0), F(1.0)) * a)
    return np.stack(out, axis=-1).astype(np.float32)


def composite_rgb(d, s, mode):
    """compositeRGB (tile_fragment.inc.glsl:488-523)."""
    if mode == "multiply":
        return d * s
    if mode == "screen":
        return _screen(d, s)
    if mode == "overlay":
        return _hard_light(s, d)
    if mode == "darken":
        return np.minimum(d, s)
    if mode == "lighten":
        return np.maximum(d, s)
    if mode == "color_dodge":
        return _color_dodge(d, s)
    if mode == "color_burn":
        return F(1.0) - _color_dodge(F(1.0) - d, F(1.0) - s)
    if mode == "hard_light":
        return _hard_light(d, s)
    if mode == "soft_light":
        return _soft_light(d, s)
    if mode == "difference":
        return np.abs(d - s)
    if mode == "exclusion":
        return d + s - F(2.0) * d * s
    dh, sh = _rgb_to_hsl(d), _rgb_to_hsl(s)
    pick = {"hue": (sh, dh, dh), "saturation": (dh, sh, dh), "color": (sh, sh, dh), "luminosity": (dh, dh, sh)}[mode]
    return _hsl_to_rgb(np.stack([pick[0][..., 0], pick[1][..., 1], pick[2][..., 2]], axis=-1))


def blend(dest, color, mask, mode="src_over", drawn=None):
    """One path over the frame. dest: (h, w, 4) premultiplied float32; color: (h, w, 4) or (4,) NOT premultiplied;
    mask: (h, w) mask alpha. calculateColor (:583-614) + the mode's blend state (blend.rs).
    drawn: (h, w) bool, the pixels of the tiles the path draws (its tiles with fills or a non-zero backdrop: empty
    tiles are never drawn, builder.rs:1014-1016). Needed by the destructive modes (effects.rs:222-235), which also
    change pixels the mask leaves out; for every other mode a pixel with mask 0 keeps its value anyway."""
    if mode in DESTRUCTIVE and drawn is None:
        raise ValueError("a destructive blend mode needs the set of drawn tiles")
    dest = np.asarray(dest, np.float32)
    out = _blend_everywhere(dest, color, mask, mode)
    if drawn is not None:
        out = np.where(np.asarray(drawn, bool)[..., None], out, dest).astype(np.float32)
    return out


def _blend_everywhere(dest, color, mask, mode):
    color = np.broadcast_to(np.asarray(color, np.float32), dest.shape)
    sa = (color[..., 3] * np.asarray(mask, np.float32))[..., None]
    d_rgb, da = dest[..., :3], dest[..., 3:4]
    s_rgb = color[..., :3]
    if BLEND_MODES.index(mode) >= BLEND_MODES.index("darken"):
        blended = composite_rgb(d_rgb, s_rgb, mode)
        rgb = sa * (F(1.0) - da) * s_rgb + sa * da * blended + (F(1.0) - sa) * d_rgb
        return np.concatenate([np.clip(rgb, F(0.0), F(1.0)), np.ones_like(sa)], axis=-1).astype(np.float32)  # UNORM target
    one = np.ones_like(sa)
    zero = 0 * one
    sf, df = {"src_over": (one, one - sa), "dest_over": (one - da, one), "dest_out": (zero, one - sa),
              "src_atop": (da, one - sa), "xor": (one - da, one - sa), "lighter": (one, one),
              # the destructive modes (blend.rs:46-53,72-98,117-125; Copy: blending disabled, :144)
              "clear": (zero, zero), "copy": (one, zero), "src_in": (da, zero), "src_out": (one - da, zero),
              "dest_in": (zero, sa), "dest_atop": (one - da, sa)}[mode]
    src = np.concatenate([s_rgb * sa, sa], axis=-1)
    out = src * sf + dest * df
    return np.clip(out, F(0.0), F(1.0)).astype(np.float32)  # the render target is UNORM (Lighter adds; a colour matrix can overshoot)
