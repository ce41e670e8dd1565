"""Generates Pathfinder's area lookup table (`textures/area-lut.png`) from its definition.

The reference ships the table as a PNG resource produced by `utils/area-lut/src/main.rs:29-96`;
the renderer loads it through its ResourceLoader (`renderer/src/gpu/renderer.rs:207-214`). This
module restates that generator in float32 so the CUDA backend does not depend on the reference's
resource directory. tests/test_area_lut.py checks the result byte-for-byte against the PNG
whenever the reference checkout is present, and against a committed checksum otherwise.

Texel (u, v), channel k holds round(255 * area(y - k, dydx)) with y = (u - 128) / 16 and
dydx = -v / 16: the area of the unit pixel centred on the origin that lies below the line through
(0, y) with slope dydx ... columns 0 and 255 are forced to 255 / 0.
"""
from __future__ import annotations

import numpy as np

WIDTH = 256
HEIGHT = 256
SHA256 = "2352b0ec6b5ba601bb46ac5a3b1ae6106903670ae7a80add38f7ab2ecb9856fc"  # of the (256, 256, 4) uint8 bytes

_f = np.float32


def _solve_line_y(p0x, p0y, p1x, p1y, y):
    # solve_line_y (main.rs:16-19): Point2D::new(p0.x - (p0.y - y) / m, y)
    with np.errstate(divide="ignore", invalid="ignore"):
        m = (p1y - p0y) / (p1x - p0x)
        return p0x - (p0y - y) / m, np.full_like(p0x, y)


def _area_tri(p0x, p0y, p1x, p1y):
    return _f(0.5) * (p1x - p0x) * (p0y - p1y)  # main.rs:21-23


def _area_rect(p0x, p0y, p1x, p1y):
    return (p1x - p0x) * (p0y - p1y)  # main.rs:25-27


def _area(y: np.ndarray, dydx: np.ndarray) -> np.ndarray:
    """area() (main.rs:29-68), vectorised over float32 arrays."""
    x_left, x_right = _f(-0.5), _f(0.5)
    y_left = dydx * x_left + y
    y_right = dydx * x_right + y
    p0x, p0y = np.full_like(y, x_left), y_left
    p1x, p1y = np.full_like(y, x_right), y_right
    p2x, p2y = _solve_line_y(p0x, p0y, p1x, p1y, _f(-0.5))
    p3x, p3y = p1x, np.full_like(y, _f(-0.5))
    p4x, p4y = _solve_line_y(p0x, p0y, p1x, p1y, _f(0.5))
    p7x, p7y = p1x, np.full_like(y, _f(0.5))
    with np.errstate(invalid="ignore"):
        tri01 = _area_tri(p0x, p0y, p1x, p1y)
        tri21 = _area_tri(p2x, p2y, p1x, p1y)
        rect07 = _area_rect(p0x, p0y, p7x, p7y)
        tri04 = _area_tri(p0x, p0y, p4x, p4y)
        rect03 = _area_rect(p0x, p0y, p3x, p3y)
        case0 = tri01 - tri21 - rect07 + tri04
        case6 = tri01 - rect07 + tri04
        case1 = tri01 - tri21 - rect07
        case4 = tri01 - rect07
        case2 = -rect07 + rect03
    out = np.where(
        p0y > _f(0.5),
        np.where(p1y < _f(-0.5), case0, np.where(p1y < _f(0.5), case6, _f(0.0))),
        np.where(p0y > _f(-0.5), np.where(p1y < _f(-0.5), case1, case4), case2),
    )
    return out.astype(np.float32)


def generate() -> np.ndarray:
    """Returns the 256x256 RGBA8 table as a (256, 256, 4) uint8 array indexed [v][u][k]."""
    u = np.arange(WIDTH, dtype=np.float32)[None, :].repeat(HEIGHT, axis=0)
    v = np.arange(HEIGHT, dtype=np.float32)[:, None].repeat(WIDTH, axis=1)
    y = (u - _f(WIDTH // 2)) / _f(16.0)
    dydx = -v / _f(16.0)
    out = np.zeros((HEIGHT, WIDTH, 4), dtype=np.uint8)
    for k in range(4):
        a = _area(y - _f(k), dydx) * _f(255.0)
        # f32::round (half away from zero) then `as u8` (saturating).
        r = np.where(a >= 0, np.floor(a + _f(0.5)), -np.floor(-a + _f(0.5)))
        out[:, :, k] = np.clip(np.nan_to_num(r, nan=0.0), 0, 255).astype(np.uint8)
    out[:, 0, :] = 255      # main.rs:78-80
    out[:, WIDTH - 1, :] = 0  # main.rs:81-83
    return out
