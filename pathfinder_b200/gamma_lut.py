"""Generates Pathfinder's gamma-correction lookup table (`textures/gamma-lut.png`) from its definition.

The reference ships the table as a PNG resource produced by `utils/gamma-lut/src/main.rs:39-58` (a port of Skia's
preblend tables: `utils/gamma-lut/src/gamma_lut.rs:196-290`, CONTRAST = 0, paint and device gamma 0 = sRGB); the
text filter samples it as `texture(gammaLUT, vec2(alpha, 1 - bgColor))` (`shaders/tile_fragment.inc.glsl:122-132`).
This module restates the generator in float32 so that the text path does not depend on the
reference's resource directory. tests/test_gamma_lut.py checks the result byte-for-byte against the PNG whenever the
reference checkout is present, against a committed checksum otherwise, and restates the reference's own unit test
(`gamma_lut.rs:316-353`).

Row i (of 8) is the table for source luminance scale255(3, i); entry a is the coverage to blend with so that an OVER
blend of the sRGB source onto the "perceptual inverse" destination 1 - src gives the linear-light result.
"""
from __future__ import annotations

import numpy as np

LUM_BITS = 3                 # gamma_lut.rs:115
TABLES = 1 << LUM_BITS
WIDTH = 256
SHA256 = "eedfd96d81bec63e3f92e2e381dcf5ea9aa537174229bb821eb3e104825ef118"  # of the (8, 256) uint8 bytes

_f = np.float32


def scale255(n: int, base: int) -> int:
    """gamma_lut.rs:87-98: scales base <= 2^n - 1 to 0..255 by bit replication (u8 arithmetic)."""
    base = (base << (8 - n)) & 0xFF
    lum, i = base, n
    while i < 8:
        lum |= base >> i
        i += n
    return lum


def srgb_to_luma(luminance):
    """LuminanceColorSpace::Srgb.to_luma (gamma_lut.rs:47-56), float32."""
    luminance = np.asarray(luminance, dtype=_f)
    with np.errstate(invalid="ignore"):
        hi = np.power((luminance + _f(0.055)) / _f(1.055), _f(2.4), dtype=_f)
    return np.where(luminance <= _f(0.04045), luminance / _f(12.92), hi).astype(_f)


def srgb_from_luma(luma):
    """LuminanceColorSpace::Srgb.from_luma (gamma_lut.rs:60-72), float32."""
    luma = np.asarray(luma, dtype=_f)
    with np.errstate(invalid="ignore"):
        hi = _f(1.055) * np.power(luma, _f(1.0) / _f(2.4), dtype=_f) - _f(0.055)
    return np.where(luma <= _f(0.0031308), luma * _f(12.92), hi).astype(_f)


def gamma_to_luma(gamma):
    g = _f(gamma)
    return (lambda x: np.power(np.asarray(x, dtype=_f), g, dtype=_f)), \
           (lambda x: np.power(np.asarray(x, dtype=_f), _f(1.0) / g, dtype=_f))


def build_gamma_correcting_lut(src: int, contrast: float = 0.0, to_luma=srgb_to_luma, from_luma=srgb_from_luma,
                               src_to_luma=None) -> np.ndarray:
    """build_gamma_correcting_lut (gamma_lut.rs:196-257) for one source luminance; returns 256 u8 entries."""
    src_to_luma = src_to_luma or to_luma
    src_f = _f(src) / _f(255.0)
    lin_src = src_to_luma(src_f)
    dst = _f(1.0) - src_f                      # the guess at the destination: the perceptual inverse
    lin_dst = to_luma(dst)
    adjusted_contrast = _f(contrast) * lin_dst
    raw_srca = np.arange(256, dtype=_f) / _f(255.0)
    srca = raw_srca + ((_f(1.0) - raw_srca) * adjusted_contrast * raw_srca)  # apply_contrast
    if abs(src_f - dst) < _f(1.0 / 256.0):
        result = srca
    else:
        dsta = _f(1.0) - srca
        lin_out = lin_src * srca + dsta * lin_dst
        out = from_luma(lin_out)
        result = (out - dst) / (src_f - dst)   # undo what the OVER blend will do
    v = np.floor(_f(255.0) * result + _f(0.5)).astype(np.int32)  # round_to_u8
    assert ((0 <= v) & (v < 256)).all()
    return v.astype(np.uint8)


def generate() -> np.ndarray:
    """GammaLut::new(0, 0, 0).tables (gamma_lut.rs:259-291; main.rs:39-47): (8, 256) uint8."""
    return np.stack([build_gamma_correcting_lut(scale255(LUM_BITS, i)) for i in range(TABLES)])
