"""ctypes wrapper of the CPU oracle (oracle/libpf_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. It must never be imported from pathfinder_b200/.

PARITY UNPINNED: see oracle/pf_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpf_oracle.so")


def build(force: bool = False) -> str:
    """Compiles the oracle with its Makefile (gcc only)."""
    src = os.path.join(_HERE, "pf_oracle.cpp")
    hdr = os.path.join(_HERE, "pf_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpf_oracle.so", "CXX=g++"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Scene(C.Structure):
    _fields_ = [
        ("points", C.c_void_p), ("point_flags", C.c_void_p), ("contour_offsets", C.c_void_p),
        ("n_points", C.c_uint32), ("n_contours", C.c_uint32),
        ("n_clip_paths", C.c_uint32), ("clip_contour_ranges", C.c_void_p),
        ("clip_fill_rules", C.c_void_p),
        ("n_draw_paths", C.c_uint32), ("draw_contour_ranges", C.c_void_p),
        ("draw_fill_rules", C.c_void_p), ("draw_paints", C.c_void_p),
        ("draw_clip_paths", C.c_void_p),
        ("n_paints", C.c_uint32), ("paint_colors", C.c_void_p),
        ("view_box", C.c_float * 4),
    ]


class _Options(C.Structure):
    _fields_ = [
        ("has_transform", C.c_int32), ("transform", C.c_float * 6), ("dilation", C.c_float * 2),
        ("subpixel_aa_enabled", C.c_int32), ("strip_tile_y0", C.c_int32),
        ("strip_tile_y1", C.c_int32),
    ]


FILL_DTYPE = np.dtype([("from_x", "<u2"), ("from_y", "<u2"), ("to_x", "<u2"), ("to_y", "<u2"),
                       ("link", "<u4")])
TILE_DTYPE = np.dtype([("tile_x", "<i2"), ("tile_y", "<i2"), ("alpha_tile_id", "<u4"),
                       ("path_id", "<u4"), ("color", "<u2"), ("ctrl", "u1"), ("backdrop", "i1")])
CLIP_DTYPE = np.dtype([("dest_tile_id", "<u4"), ("dest_backdrop", "<i4"), ("src_tile_id", "<u4"),
                       ("src_backdrop", "<i4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_LIB_PATH)
        l.pfo_build.restype = C.c_void_p
        l.pfo_build.argtypes = [C.POINTER(_Scene), C.POINTER(_Options), C.c_int]
        l.pfo_built_destroy.argtypes = [C.c_void_p]
        for name, res in [("pfo_fill_count", C.c_size_t), ("pfo_fills", C.c_void_p),
                          ("pfo_fill_path_offsets", C.c_void_p), ("pfo_tile_count", C.c_size_t),
                          ("pfo_tiles", C.c_void_p), ("pfo_clip_count", C.c_size_t),
                          ("pfo_clips", C.c_void_p), ("pfo_alpha_tile_count", C.c_uint32),
                          ("pfo_line_segment_count", C.c_uint64),
                          ("pfo_input_segment_count", C.c_uint64),
                          ("pfo_bbox_tile_count", C.c_uint64), ("pfo_build_seconds", C.c_double),
                          ("pfo_build_seconds_paths", C.c_double)]:
            f = getattr(l, name)
            f.restype = res
            f.argtypes = [C.c_void_p]
        l.pfo_z_buffer.restype = C.c_void_p
        l.pfo_z_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_int32 * 4)]
        l.pfo_path_lines.restype = C.c_size_t
        l.pfo_path_lines.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        l.pfo_set_keep_lines.argtypes = [C.c_int]
        l.pfo_alpha_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.pfo_render.argtypes = [C.c_void_p, C.POINTER(_Scene), C.c_void_p, C.POINTER(C.c_float * 4),
                                 C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        l.pfo_render_crop.argtypes = [C.c_void_p, C.POINTER(_Scene), C.c_void_p, C.POINTER(C.c_float * 4)] + \
                                     [C.c_uint32] * 6 + [C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _copy(ptr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


@dataclass
class OracleScene:
    """Keeps the numpy arrays alive next to the C struct that borrows them."""
    c: _Scene
    keep: tuple


def make_scene(*, points, point_flags, contour_offsets, draw_contour_ranges, draw_fill_rules,
               draw_paints, paint_colors, view_box, draw_clip_paths=None, clip_contour_ranges=None,
               clip_fill_rules=None) -> OracleScene:
    points = _arr(points, np.float32).reshape(-1, 2)
    point_flags = _arr(point_flags, np.uint8)
    contour_offsets = _arr(contour_offsets, np.uint32)
    draw_contour_ranges = _arr(draw_contour_ranges, np.uint32).reshape(-1, 2)
    n_draw = draw_contour_ranges.shape[0]
    draw_fill_rules = _arr(draw_fill_rules, np.uint8)
    draw_paints = _arr(draw_paints, np.uint16)
    if draw_clip_paths is None:
        draw_clip_paths = np.full(n_draw, 0xffffffff, dtype=np.uint32)
    draw_clip_paths = _arr(draw_clip_paths, np.uint32)
    if clip_contour_ranges is None:
        clip_contour_ranges = np.zeros((0, 2), dtype=np.uint32)
        clip_fill_rules = np.zeros(0, dtype=np.uint8)
    clip_contour_ranges = _arr(clip_contour_ranges, np.uint32).reshape(-1, 2)
    clip_fill_rules = _arr(clip_fill_rules, np.uint8)
    paint_colors = _arr(paint_colors, np.uint8).reshape(-1, 4)
    assert len(point_flags) == len(points)
    assert len(draw_fill_rules) == n_draw and len(draw_paints) == n_draw
    assert contour_offsets[-1] == len(points)
    s = _Scene()
    s.points = points.ctypes.data
    s.point_flags = point_flags.ctypes.data
    s.contour_offsets = contour_offsets.ctypes.data
    s.n_points = len(points)
    s.n_contours = len(contour_offsets) - 1
    s.n_clip_paths = clip_contour_ranges.shape[0]
    s.clip_contour_ranges = clip_contour_ranges.ctypes.data
    s.clip_fill_rules = clip_fill_rules.ctypes.data
    s.n_draw_paths = n_draw
    s.draw_contour_ranges = draw_contour_ranges.ctypes.data
    s.draw_fill_rules = draw_fill_rules.ctypes.data
    s.draw_paints = draw_paints.ctypes.data
    s.draw_clip_paths = draw_clip_paths.ctypes.data
    s.n_paints = paint_colors.shape[0]
    s.paint_colors = paint_colors.ctypes.data
    s.view_box = (C.c_float * 4)(*[float(v) for v in view_box])
    keep = (points, point_flags, contour_offsets, draw_contour_ranges, draw_fill_rules, draw_paints,
            draw_clip_paths, clip_contour_ranges, clip_fill_rules, paint_colors)
    return OracleScene(s, keep)


def make_options(transform=None, dilation=(0.0, 0.0), subpixel_aa_enabled=False, strip=None) -> _Options:
    o = _Options()
    if transform is not None:
        o.has_transform = 1
        o.transform = (C.c_float * 6)(*[float(v) for v in transform])
    else:
        o.has_transform = 0
        o.transform = (C.c_float * 6)(1, 0, 0, 1, 0, 0)
    o.dilation = (C.c_float * 2)(float(dilation[0]), float(dilation[1]))
    o.subpixel_aa_enabled = int(bool(subpixel_aa_enabled))
    if strip is not None:
        o.strip_tile_y0, o.strip_tile_y1 = int(strip[0]), int(strip[1])
    return o


class Built:
    """Result of one oracle build (the D3D9-level command payloads)."""

    def __init__(self, scene: OracleScene, options: _Options, n_threads: int = 1, keep_lines: bool = False):
        l = lib()
        self.scene = scene
        l.pfo_set_keep_lines(int(keep_lines))
        self._h = l.pfo_build(C.byref(scene.c), C.byref(options), n_threads)
        l.pfo_set_keep_lines(0)
        h = self._h
        self.fills = _copy(l.pfo_fills(h), l.pfo_fill_count(h), FILL_DTYPE)
        npaths = scene.c.n_clip_paths + scene.c.n_draw_paths
        self.fill_path_offsets = _copy(l.pfo_fill_path_offsets(h), npaths + 1, np.uint32)
        self.tiles = _copy(l.pfo_tiles(h), l.pfo_tile_count(h), TILE_DTYPE)
        self.clips = _copy(l.pfo_clips(h), l.pfo_clip_count(h), CLIP_DTYPE)
        rect = (C.c_int32 * 4)()
        zptr = l.pfo_z_buffer(h, C.byref(rect))
        self.z_rect = tuple(rect)
        w, hh = max(rect[2] - rect[0], 0), max(rect[3] - rect[1], 0)
        self.z_buffer = _copy(zptr, w * hh, np.int32).reshape(hh, w)
        self.alpha_tile_count = int(l.pfo_alpha_tile_count(h))
        self.line_segment_count = int(l.pfo_line_segment_count(h))
        self.input_segment_count = int(l.pfo_input_segment_count(h))
        self.bbox_tile_count = int(l.pfo_bbox_tile_count(h))
        self.build_seconds = float(l.pfo_build_seconds(h))

    def path_lines(self, path: int) -> np.ndarray:
        l = lib()
        n = l.pfo_path_lines(self._h, path, None, 0)
        out = np.zeros((n, 4), dtype=np.float32)
        if n:
            l.pfo_path_lines(self._h, path, out.ctypes.data, n)
        return out

    def alpha_masks(self, area_lut: np.ndarray) -> np.ndarray:
        lut = _arr(area_lut, np.uint8).reshape(256, 256, 4)
        out = np.zeros((self.alpha_tile_count, 16, 16), dtype=np.float32)
        lib().pfo_alpha_masks(self._h, lut.ctypes.data, out.ctypes.data)
        return out

    def render(self, area_lut: np.ndarray, width: int, height: int, background=(0, 0, 0, 0),
               want_f32: bool = False):
        lut = _arr(area_lut, np.uint8).reshape(256, 256, 4)
        out = np.zeros((height, width, 4), dtype=np.uint8)
        outf = np.zeros((height, width, 4), dtype=np.float32) if want_f32 else None
        bg = (C.c_float * 4)(*[float(v) for v in background])
        lib().pfo_render(self._h, C.byref(self.scene.c), lut.ctypes.data, C.byref(bg), width, height,
                         out.ctypes.data, outf.ctypes.data if want_f32 else None)
        return (out, outf) if want_f32 else out

    def render_crop(self, area_lut: np.ndarray, frame_size, origin, size, background=(0, 0, 0, 0)):
        """RGBA8 of the pixels [x0, x0 + w) x [y0, y0 + h) of a frame of frame_size = (W, H) pixels; only the
        alpha masks that reach the crop are evaluated."""
        lut = _arr(area_lut, np.uint8).reshape(256, 256, 4)
        (x0, y0), (w, h) = origin, size
        out = np.zeros((h, w, 4), dtype=np.uint8)
        bg = (C.c_float * 4)(*[float(v) for v in background])
        lib().pfo_render_crop(self._h, C.byref(self.scene.c), lut.ctypes.data, C.byref(bg), int(frame_size[0]),
                              int(frame_size[1]), int(x0), int(y0), int(w), int(h), out.ctypes.data, None)
        return out

    def close(self):
        if self._h:
            lib().pfo_built_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def time_build(scene: OracleScene, options: _Options, n_threads: int) -> float:
    """Runs one build and returns its wall seconds (the reference's cpu_build_time analogue)."""
    l = lib()
    h = l.pfo_build(C.byref(scene.c), C.byref(options), n_threads)
    s = float(l.pfo_build_seconds(h))
    l.pfo_built_destroy(h)
    return s
