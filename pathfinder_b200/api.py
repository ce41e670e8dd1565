"""Python mirror of the `pathfinder_renderer` surface for the CUDA backend, over the C ABI.

Names and call protocol follow the reference so parity tests read like its own drivers:
  Scene / BuildOptions                    renderer/src/scene.rs:36-226, options.rs:50-61
  Scene::build / build_and_render         renderer/src/scene.rs:290-297, 369-378
  Renderer::{begin_scene, render_command, end_scene}   renderer/src/gpu/renderer.rs:350-460
  RendererOptions / RendererMode / RendererLevel       renderer/src/gpu/options.rs:19-119
All compute happens in libpf_cuda.so; nothing here has a CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import area_lut as _area_lut
from .flat_scene import FlatScene

FILL_DTYPE = np.dtype([("from_x", "<u2"), ("from_y", "<u2"), ("to_x", "<u2"), ("to_y", "<u2"), ("link", "<u4")])
TILE_DTYPE = np.dtype([("tile_x", "<i2"), ("tile_y", "<i2"), ("alpha_tile_id", "<u4"), ("path_id", "<u4"),
                       ("color", "<u2"), ("ctrl", "u1"), ("backdrop", "i1")])
BLEND_MODE_SRC_OVER = 4  # PF_BLEND_MODE_SRC_OVER (BlendMode::SrcOver, content/src/effects.rs:99-163)
# BlendMode, numbered like include/pf_cuda.h PF_BLEND_MODE_*
BLEND_MODES = {name: i for i, name in enumerate(
    ["clear", "copy", "src_in", "src_out", "src_over", "src_atop", "dest_in", "dest_out", "dest_over", "dest_atop", "xor",
     "lighter", "darken", "lighten", "multiply", "screen", "hard_light", "overlay", "color_dodge", "color_burn",
     "soft_light", "difference", "exclusion", "hue", "saturation", "color", "luminosity"])}
CLIP_DTYPE = np.dtype([("dest_tile_id", "<u4"), ("dest_backdrop", "<i4"), ("src_tile_id", "<u4"), ("src_backdrop", "<i4")])


class RendererLevel:
    D3D9 = L.PF_RENDERER_LEVEL_D3D9
    D3D11 = L.PF_RENDERER_LEVEL_D3D11


class Transform2F:
    """geometry/src/transform2d.rs Transform2F: matrix [m11 m12; m21 m22] + vector."""

    def __init__(self, m11=1.0, m12=0.0, m21=0.0, m22=1.0, tx=0.0, ty=0.0):
        f = np.float32
        self.m11, self.m12, self.m21, self.m22, self.tx, self.ty = f(m11), f(m12), f(m21), f(m22), f(tx), f(ty)

    @staticmethod
    def from_scale(sx, sy=None) -> "Transform2F":
        return Transform2F(sx, 0, 0, sx if sy is None else sy, 0, 0)

    def translate(self, tx, ty) -> "Transform2F":
        """Transform2F::translate: from_translation(v) * self (transform2d.rs:262-266)."""
        f = np.float32
        return Transform2F(self.m11, self.m12, self.m21, self.m22, f(self.tx) + f(tx), f(self.ty) + f(ty))

    def as_oracle_tuple(self):
        return (self.m11, self.m21, self.m12, self.m22, self.tx, self.ty)

    def _c(self) -> L.PFTransform2F:
        return L.PFTransform2F(L.PFMatrix2x2F(self.m11, self.m12, self.m21, self.m22), L.PFVector2F(self.tx, self.ty))


class BuildOptions:
    """renderer/src/options.rs:50-61."""

    def __init__(self, transform: Transform2F | None = None, dilation=(0.0, 0.0), subpixel_aa_enabled=False):
        self.transform = transform
        self.dilation = dilation
        self.subpixel_aa_enabled = subpixel_aa_enabled
        lib = L.lib()
        self._h = lib.PFBuildOptionsCreate()
        if transform is not None:
            t = transform._c()
            lib.PFBuildOptionsSetTransform(self._h, lib.PFRenderTransformCreate2D(C.byref(t)))
        d = L.PFVector2F(float(dilation[0]), float(dilation[1]))
        lib.PFBuildOptionsSetDilation(self._h, C.byref(d))
        lib.PFBuildOptionsSetSubpixelAAEnabled(self._h, int(bool(subpixel_aa_enabled)))

    def __del__(self):
        try:
            if self._h:
                L.lib().PFBuildOptionsDestroy(self._h)
                self._h = None
        except Exception:
            pass


class Scene:
    """renderer/src/scene.rs Scene (solid-colour draw paths)."""

    def __init__(self):
        self._h = L.lib().PFSceneCreate()

    @staticmethod
    def from_flat(flat: FlatScene) -> "Scene":
        s = Scene()
        s.set_view_box(flat.view_box)
        s.push_flat(flat)
        return s

    def push_flat(self, flat: FlatScene):
        """Appends every clip and draw path of a FlatScene (its view box is not touched)."""
        s = self
        lib = L.lib()
        remap = np.zeros(len(flat.paint_colors), dtype=np.uint16)
        for i, c in enumerate(flat.paint_colors):
            remap[i] = s.push_paint(c)
        paints = np.ascontiguousarray(remap[flat.paints], dtype=np.uint16)
        for (c0, c1), rule in zip(flat.clip_contour_ranges, flat.clip_fill_rules):
            p0, p1 = int(flat.contour_offsets[c0]), int(flat.contour_offsets[c1])
            s.push_clip_path(flat.points[p0:p1], flat.point_flags[p0:p1], flat.contour_offsets[c0:c1 + 1] - p0, int(rule))
        n_draw_contours = int(flat.path_contour_offsets[-1])
        n_draw_points = int(flat.contour_offsets[n_draw_contours])
        clip_ids = flat.draw_clip_paths if flat.n_clip_paths else None
        L.check(lib.PFScenePushDrawPaths(
            s._h, flat.points.ctypes.data, flat.point_flags.ctypes.data, n_draw_points,
            flat.contour_offsets.ctypes.data, n_draw_contours, flat.path_contour_offsets.ctypes.data,
            flat.n_paths, paints.ctypes.data, flat.fill_rules.ctypes.data,
            clip_ids.ctypes.data if clip_ids is not None else None))

    def set_view_box(self, view_box):
        r = L.PFRectF(L.PFVector2F(view_box[0], view_box[1]), L.PFVector2F(view_box[2], view_box[3]))
        L.lib().PFSceneSetViewBox(self._h, C.byref(r))

    def view_box(self):
        r = L.PFRectF()
        L.lib().PFSceneGetViewBox(self._h, C.byref(r))
        return (r.origin.x, r.origin.y, r.lower_right.x, r.lower_right.y)

    def push_paint(self, rgba) -> int:
        c = L.PFColorU(int(rgba[0]), int(rgba[1]), int(rgba[2]), int(rgba[3]))
        return int(L.lib().PFScenePushPaint(self._h, C.byref(c)))

    def push_render_target(self, width: int, height: int) -> int:
        """Scene::push_render_target: the paths pushed until pop_render_target are drawn into an off-screen image."""
        rid = int(L.lib().PFScenePushRenderTarget(self._h, int(width), int(height)))
        if rid == 0xFFFFFFFF:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, L.lib().PFCudaGetLastError().decode())
        return rid

    def pop_render_target(self):
        L.lib().PFScenePopRenderTarget(self._h)

    def push_render_target_pattern(self, render_target_id: int, transform=None, text_filter=None, pattern_filter=None) -> int:
        """Paint::from_pattern(Pattern::from_render_target(id, size)). transform = (m11, m12, m21, m22, tx, ty) maps
        render-target pixels to scene coordinates (pattern.apply_transform). text_filter = dict(fg=rgb, bg=rgb,
        kernel=(4 floats) | None, gamma=bool) for PatternFilter::Text."""
        t = None
        if transform is not None:
            m11, m12, m21, m22, tx, ty = [float(v) for v in transform]
            t = L.PFTransform2F(L.PFMatrix2x2F(m11, m12, m21, m22), L.PFVector2F(tx, ty))
        f = self._pattern_filter_c(pattern_filter)
        if text_filter is not None:
            f = L.PFFilter()
            f.kind = L.PF_FILTER_TEXT
            fg, bg = list(text_filter["fg"]) + [1.0], list(text_filter["bg"]) + [1.0]
            for i in range(4):
                f.params[i], f.params[4 + i] = float(fg[i]), float(bg[i])
            if text_filter.get("kernel") is not None:
                f.flags |= L.PF_FILTER_FLAG_TEXT_HAS_KERNEL
                for i in range(4):
                    f.params[8 + i] = float(text_filter["kernel"][i])
            if text_filter.get("gamma"):
                f.flags |= L.PF_FILTER_FLAG_TEXT_GAMMA_CORRECTION
        pid = int(L.lib().PFScenePushPaintRenderTargetPattern(self._h, int(render_target_id),
                                                             C.byref(t) if t is not None else None,
                                                             C.byref(f) if f is not None else None))
        if pid == 0xFFFF:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, L.lib().PFCudaGetLastError().decode())
        return pid

    @staticmethod
    def _transform_c(transform):
        if transform is None:
            return None
        m11, m12, m21, m22, tx, ty = [float(v) for v in transform]
        return L.PFTransform2F(L.PFMatrix2x2F(m11, m12, m21, m22), L.PFVector2F(tx, ty))

    @staticmethod
    def _pattern_filter_c(pattern_filter):
        """dict(blur=sigma, direction="x" | "y") or dict(color_matrix=20 floats: the five F32x4 columns)."""
        if pattern_filter is None:
            return None
        f = L.PFFilter()
        if "blur" in pattern_filter:
            f.kind = L.PF_FILTER_BLUR
            f.params[0] = float(pattern_filter["blur"])
            if pattern_filter.get("direction", "x") == "y":
                f.flags |= L.PF_FILTER_FLAG_BLUR_Y
        elif "color_matrix" in pattern_filter:
            f.kind = L.PF_FILTER_COLOR_MATRIX
            for i, v in enumerate(pattern_filter["color_matrix"]):
                f.params[i] = float(v)
        else:
            raise ValueError("unknown pattern filter")
        return f

    def push_image_pattern(self, pixels, transform=None, repeat_x=False, repeat_y=False, smoothing=True,
                           pattern_filter=None) -> int:
        """Paint::from_pattern(Pattern::from_image(image)): pixels = (h, w, 4) uint8, not premultiplied; transform maps
        image pixels to scene coordinates."""
        px = np.ascontiguousarray(pixels, dtype=np.uint8)
        h, w = px.shape[:2]
        flags = (L.PF_PATTERN_FLAG_REPEAT_X if repeat_x else 0) | (L.PF_PATTERN_FLAG_REPEAT_Y if repeat_y else 0) | \
                (0 if smoothing else L.PF_PATTERN_FLAG_NO_SMOOTHING)
        t, f = self._transform_c(transform), self._pattern_filter_c(pattern_filter)
        pid = int(L.lib().PFScenePushPaintImagePattern(self._h, px.ctypes.data, w, h, C.byref(t) if t is not None else None,
                                                       flags, C.byref(f) if f is not None else None))
        if pid == 0xFFFF:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, L.lib().PFCudaGetLastError().decode())
        return pid

    def push_gradient(self, stops, line, radii=None, transform=None, repeat=False) -> int:
        """Paint::from_gradient: stops = [(offset, (r, g, b, a) bytes)], sorted; line = ((x0, y0), (x1, y1));
        radii = (r0, r1) makes it radial, in the space `transform` maps to scene coordinates."""
        g = L.PFGradient()
        g.kind = L.PF_GRADIENT_LINEAR if radii is None else L.PF_GRADIENT_RADIAL
        g.wrap = L.PF_GRADIENT_WRAP_REPEAT if repeat else L.PF_GRADIENT_WRAP_CLAMP
        g.from_ = L.PFVector2F(float(line[0][0]), float(line[0][1]))
        g.to = L.PFVector2F(float(line[1][0]), float(line[1][1]))
        if radii is not None:
            g.radii[0], g.radii[1] = float(radii[0]), float(radii[1])
        t = self._transform_c(transform if transform is not None else (1, 0, 0, 1, 0, 0))
        g.transform = t
        arr = (L.PFColorStop * len(stops))()
        for i, (offset, rgba) in enumerate(stops):
            arr[i].offset = float(offset)
            arr[i].color = L.PFColorU(int(rgba[0]), int(rgba[1]), int(rgba[2]), int(rgba[3]))
        g.stops = arr
        g.stop_count = len(stops)
        pid = int(L.lib().PFScenePushPaintGradient(self._h, C.byref(g)))
        if pid == 0xFFFF:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, L.lib().PFCudaGetLastError().decode())
        return pid

    def push_draw_path(self, points, point_flags, contour_offsets, paint_id, fill_rule=0, blend_mode=BLEND_MODE_SRC_OVER,
                       clip_path_id=0xFFFFFFFF) -> int:
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
        fl = np.ascontiguousarray(point_flags, dtype=np.uint8)
        co = np.ascontiguousarray(contour_offsets, dtype=np.uint32)
        return int(L.lib().PFScenePushDrawPath(self._h, pts.ctypes.data, fl.ctypes.data, co.ctypes.data,
                                               len(co) - 1, paint_id, fill_rule, blend_mode, clip_path_id))

    def push_stroked_path(self, points, point_flags, contour_offsets, contour_closed, paint_id, line_width,
                          line_join="miter", miter_limit=10.0, line_cap="butt", clip_path_id=0xFFFFFFFF) -> int:
        """What the reference's front ends do with a stroked path (svg/src/lib.rs:204-231, canvas stroke_path):
        OutlineStrokeToFill, then a draw path with the winding rule."""
        pts, flags, offsets = stroke_to_fill(points, point_flags, contour_offsets, contour_closed, line_width,
                                             line_join, miter_limit, line_cap)
        return self.push_draw_path(pts, flags, offsets, paint_id, fill_rule=0, clip_path_id=clip_path_id)

    def push_glyph(self, font: "Font", glyph_id: int, glyph_offset, font_size: float, paint_id: int,
                   transform: "Transform2F | None" = None, clip_path_id=0xFFFFFFFF) -> int:
        """FontContext::push_glyph (text/src/lib.rs:83-146) for unhinted, filled text: one draw path per glyph,
        winding rule. Returns the DrawPathId, or None for a glyph without contours (the reference pushes an empty
        path there; it tiles to nothing)."""
        pts, flags, offsets = font.glyph_outline_at(glyph_id, glyph_offset, font_size, transform)
        if len(pts) == 0:
            return None
        return self.push_draw_path(pts, flags, offsets, paint_id, fill_rule=0, clip_path_id=clip_path_id)

    def push_text(self, font: "Font", text: str, origin, font_size: float, paint_id: int,
                  transform: "Transform2F | None" = None) -> float:
        """FontContext::push_text (text/src/lib.rs:197-207) with the simplest possible layout in place of skribo's
        (a dependency that is not vendored): glyphs of one font set left to right by their advances, no kerning or
        shaping. Returns the pen's x position after the last glyph."""
        f = np.float32
        x, y = f(origin[0]), f(origin[1])
        scale = f(font_size) / f(font.units_per_em)
        for ch in text:
            glyph = font.glyph_for_char(ch)
            self.push_glyph(font, glyph, (x, y), font_size, paint_id, transform)
            x = f(x + f(font.advance(glyph)) * scale)
        return float(x)

    def push_clip_path(self, points, point_flags, contour_offsets, fill_rule=0, clip_path_id=0xFFFFFFFF) -> int:
        """Scene::push_clip_path (scene.rs:99-106); returns the ClipPathId."""
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
        fl = np.ascontiguousarray(point_flags, dtype=np.uint8)
        co = np.ascontiguousarray(contour_offsets, dtype=np.uint32)
        return int(L.lib().PFScenePushClipPath(self._h, pts.ctypes.data, fl.ctypes.data, co.ctypes.data,
                                               len(co) - 1, fill_rule, clip_path_id))

    def draw_path_count(self) -> int:
        return int(L.lib().PFSceneGetDrawPathCount(self._h))

    def build(self, options: BuildOptions, listener, sink_state: L.PFSceneSinkState | None = None, strip=None):
        """Scene::build at the D3D11 level: calls listener(PFRenderCommand) per command. strip = (tile_y0, tile_y1):
        build for a renderer that owns those tile rows (PFSceneBuildForStrip)."""
        state = sink_state if sink_state is not None else L.PFSceneSinkState()
        errors = []

        def trampoline(cmd_ptr, _userdata):
            try:
                listener(cmd_ptr.contents)
                return 0
            except L.PathfinderCudaError as e:  # propagate the renderer's status
                errors.append(e)
                return e.status
            except Exception as e:  # noqa: BLE001
                errors.append(e)
                return L.PF_CUDA_ERROR_INVALID_ARGUMENT

        cb = L.LISTENER_FN(trampoline)
        if strip is not None:
            status = L.lib().PFSceneBuildForStrip(self._h, options._h, C.byref(state), cb, None, int(strip[0]), int(strip[1]))
        else:
            status = L.lib().PFSceneBuild(self._h, options._h, C.byref(state), cb, None)
        if errors:
            raise errors[0]
        L.check(status)

    def build_and_render(self, renderer: "CudaRenderer", options: BuildOptions):
        """Scene::build_and_render (scene.rs:369-378)."""
        L.check(L.lib().PFSceneBuildAndRenderCuda(self._h, renderer._h, options._h))

    def __del__(self):
        try:
            if self._h:
                L.lib().PFSceneDestroy(self._h)
                self._h = None
        except Exception:
            pass


class SceneProxy:
    """renderer/src/concurrent/scene_proxy.rs SceneProxy: the scene lives on a worker thread of the library; every call
    but the render calls returns at once."""

    def __init__(self, scene: Scene):
        """SceneProxy::from_scene: takes the scene over (the Scene object is empty afterwards)."""
        self._h = L.lib().PFSceneProxyCreateFromScene(scene._h)
        if not self._h:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, L.lib().PFCudaGetLastError().decode("utf-8", "replace"))
        scene._h = None

    def replace_scene(self, scene: Scene):
        L.check(L.lib().PFSceneProxyReplaceScene(self._h, scene._h))
        scene._h = None

    def set_view_box(self, view_box):
        r = L.PFRectF(L.PFVector2F(view_box[0], view_box[1]), L.PFVector2F(view_box[2], view_box[3]))
        L.check(L.lib().PFSceneProxySetViewBox(self._h, C.byref(r)))

    def build(self, options: BuildOptions):
        L.check(L.lib().PFSceneProxyBuild(self._h, options._h))

    def receive(self, listener):
        """The commands of the oldest queued build, to listener(PFRenderCommand), up to and including Finish."""
        errors = []

        def trampoline(cmd_ptr, _userdata):
            try:
                listener(cmd_ptr.contents)
                return 0
            except L.PathfinderCudaError as e:
                errors.append(e)
                return e.status
            except Exception as e:  # noqa: BLE001
                errors.append(e)
                return L.PF_CUDA_ERROR_INVALID_ARGUMENT

        status = L.lib().PFSceneProxyReceive(self._h, L.LISTENER_FN(trampoline), None)
        if errors:
            raise errors[0]
        L.check(status)

    def render(self, renderer: "CudaRenderer"):
        L.check(L.lib().PFSceneProxyRenderCuda(self._h, renderer._h))

    def build_and_render(self, renderer: "CudaRenderer", options: BuildOptions):
        L.check(L.lib().PFSceneProxyBuildAndRenderCuda(self._h, renderer._h, options._h))

    def copy_scene(self) -> Scene:
        h = L.lib().PFSceneProxyCopyScene(self._h)
        if not h:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_PROTOCOL, L.lib().PFCudaGetLastError().decode("utf-8", "replace"))
        s = Scene.__new__(Scene)
        s._h = h
        return s

    def close(self):
        if self._h:
            L.lib().PFSceneProxyDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


LINE_CAP = {"butt": 0, "square": 1, "round": 2}
LINE_JOIN = {"miter": 0, "bevel": 1, "round": 2}


def stroke_to_fill(points, point_flags, contour_offsets, contour_closed, line_width, line_join="miter",
                   miter_limit=10.0, line_cap="butt"):
    """OutlineStrokeToFill (content/src/stroke.rs:88-131) on flat arrays; returns (points, point_flags,
    contour_offsets) of the stroked outline, whose contours are all closed."""
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
    fl = np.ascontiguousarray(point_flags, dtype=np.uint8)
    co = np.ascontiguousarray(contour_offsets, dtype=np.uint32)
    cl = np.ascontiguousarray(contour_closed, dtype=np.uint8)
    style = L.PFStrokeStyle(float(line_width), LINE_CAP[line_cap], LINE_JOIN[line_join], float(miter_limit))
    lib = L.lib()
    h = lib.PFOutlineStrokeToFill(pts.ctypes.data, fl.ctypes.data, co.ctypes.data, cl.ctypes.data, len(co) - 1, C.byref(style))
    if not h:
        raise L.PathfinderCudaError(L.PF_CUDA_ERROR_UNSUPPORTED, lib.PFCudaGetLastError().decode("utf-8", "replace"))
    try:
        n, k = int(lib.PFOutlineGetPointCount(h)), int(lib.PFOutlineGetContourCount(h))
        out_p, out_f, out_c = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8), np.zeros(k + 1, np.uint32)
        lib.PFOutlineCopy(h, out_p.ctypes.data, out_f.ctypes.data, out_c.ctypes.data)
    finally:
        lib.PFOutlineDestroy(h)
    return out_p, out_f, out_c


def dilate_outline(points, contour_offsets, amount):
    """Outline::dilate (content/src/outline.rs:243-249) on flat arrays; returns the moved points."""
    pts = np.array(points, dtype=np.float32).reshape(-1, 2)
    co = np.ascontiguousarray(contour_offsets, dtype=np.uint32)
    a = L.PFVector2F(float(amount[0]), float(amount[1]))
    L.lib().PFOutlineDilate(pts.ctypes.data, co.ctypes.data, len(co) - 1, C.byref(a))
    return pts


def svg_path_to_outline(path_data: str):
    """SVG path data -> (points, point_flags, contour_offsets, contour_closed), see PFSvgPathDataToOutline."""
    lib = L.lib()
    h = lib.PFSvgPathDataToOutline(path_data.encode("utf-8"))
    if not h:
        raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, lib.PFCudaGetLastError().decode("utf-8", "replace"))
    try:
        n, k = int(lib.PFOutlineGetPointCount(h)), int(lib.PFOutlineGetContourCount(h))
        pts, flags = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
        offsets, closed = np.zeros(k + 1, np.uint32), np.zeros(k, np.uint8)
        lib.PFOutlineCopy(h, pts.ctypes.data, flags.ctypes.data, offsets.ctypes.data)
        lib.PFOutlineCopyClosed(h, closed.ctypes.data)
    finally:
        lib.PFOutlineDestroy(h)
    return pts, flags, offsets, closed


class Font:
    """A TrueType font read by csrc/font.cpp (the reference uses font-kit's Loader, text/src/lib.rs:80-160)."""

    def __init__(self, data: bytes):
        lib = L.lib()
        self._h = lib.PFFontCreateFromBytes(data, len(data))
        if not self._h:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, lib.PFCudaGetLastError().decode("utf-8", "replace"))
        self.units_per_em = int(lib.PFFontGetUnitsPerEm(self._h))
        self.glyph_count = int(lib.PFFontGetGlyphCount(self._h))

    @staticmethod
    def from_path(path: str) -> "Font":
        with open(path, "rb") as f:
            return Font(f.read())

    def close(self):
        if self._h:
            L.lib().PFFontDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def glyph_for_char(self, ch) -> int:
        return int(L.lib().PFFontGetGlyphForCodepoint(self._h, ord(ch) if isinstance(ch, str) else int(ch)))

    def advance(self, glyph_id: int) -> float:
        return float(L.lib().PFFontGetGlyphAdvance(self._h, int(glyph_id)))

    def outline(self, glyph_id: int):
        """(points, point_flags, contour_offsets) in font units, y up (HintingOptions::None)."""
        lib = L.lib()
        h = lib.PFFontGetGlyphOutline(self._h, int(glyph_id))
        if not h:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_INVALID_ARGUMENT, lib.PFCudaGetLastError().decode("utf-8", "replace"))
        try:
            n, k = int(lib.PFOutlineGetPointCount(h)), int(lib.PFOutlineGetContourCount(h))
            pts, flags, offsets = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8), np.zeros(k + 1, np.uint32)
            lib.PFOutlineCopy(h, pts.ctypes.data, flags.ctypes.data, offsets.ctypes.data)
        finally:
            lib.PFOutlineDestroy(h)
        return pts, flags, offsets

    def glyph_outline_at(self, glyph_id: int, glyph_offset, font_size: float, transform: Transform2F | None = None):
        """The outline FontContext::push_glyph hands to Scene::push_draw_path for unhinted text
        (text/src/lib.rs:118-146), with its f32 operation order: the cached outline is the font-unit outline scaled by
        units_per_em, and it is transformed by
            render_options.transform * from_scale(s, -s).translate(offset) * from_scale(1 / units_per_em),
        s = font_size / units_per_em (Transform2F products: matrix * matrix, matrix * vector + vector)."""
        f = np.float32
        pts, flags, offsets = self.outline(glyph_id)
        upem = f(self.units_per_em)
        cached = pts * upem                                            # OutlinePathBuilder(from_scale(units_per_em))
        s = f(font_size) / upem
        t = transform or Transform2F()
        # from_scale(s, -s).translate(offset) = from_translation(offset) * from_scale: matrix diag(s, -s), vector offset
        gm = (s, f(0.0), f(0.0), -s)
        gv = (f(glyph_offset[0]), f(glyph_offset[1]))
        # render_transform = t * glyph transform (transform2d.rs Mul: matrix = a.m * b.m, vector = a.m * b.v + a.v)
        rm = (t.m11 * gm[0] + t.m12 * gm[2], t.m11 * gm[1] + t.m12 * gm[3],
              t.m21 * gm[0] + t.m22 * gm[2], t.m21 * gm[1] + t.m22 * gm[3])
        rv = ((t.m11 * gv[0] + t.m12 * gv[1]) + t.tx, (t.m21 * gv[0] + t.m22 * gv[1]) + t.ty)
        k = f(1.0) / upem
        fm = (rm[0] * k, rm[1] * k, rm[2] * k, rm[3] * k)              # * from_scale(1 / units_per_em): vector unchanged
        x, y = cached[:, 0], cached[:, 1]
        out = np.stack([(fm[0] * x + fm[1] * y) + rv[0], (fm[2] * x + fm[3] * y) + rv[1]], axis=1).astype(np.float32)
        return out, flags, offsets


def ipc_export(device_ptr: int):
    """(handle bytes, offset) of a device allocation, to be opened by peer processes."""
    h = (C.c_uint8 * 64)()
    off = C.c_uint64()
    L.check(L.lib().PFCudaIpcExport(device_ptr, h, C.byref(off)))
    return bytes(h), int(off.value)


def strip_of_rank(tile_rows: int, rank: int, world_size: int):
    """Tile rows [y0, y1) of `rank` out of `world_size` (PFCudaStripOfRank: as equal as whole rows allow)."""
    y0, y1 = C.c_int32(), C.c_int32()
    L.lib().PFCudaStripOfRank(tile_rows, rank, world_size, C.byref(y0), C.byref(y1))
    return int(y0.value), int(y1.value)


def gather_create_id() -> bytes:
    """The 128-byte id (ncclUniqueId) rank 0 creates and ships to the other ranks before gather_init."""
    buf = (C.c_uint8 * 128)()
    L.check(L.lib().PFCudaGatherCreateId(buf))
    return bytes(buf)


class CudaRenderer:
    """Renderer<CudaDevice> at RendererLevel::D3D11."""

    def __init__(self, dest_size, background_color=None, device_ordinal: int = 0, level: int = RendererLevel.D3D11,
                 area_lut: np.ndarray | None = None):
        lib = L.lib()
        dev = lib.PFCudaDeviceCreate(device_ordinal)
        if not dev:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_NO_DEVICE, lib.PFCudaGetLastError().decode())
        self.dest_size = (int(dest_size[0]), int(dest_size[1]))
        lut = np.ascontiguousarray(_area_lut.generate() if area_lut is None else area_lut, dtype=np.uint8)
        self._options = self._make_options(self.dest_size, background_color)
        mode = L.PFRendererMode(level)
        # textures/gamma-lut.png (256 x 8 L8): only the text filter's gamma correction reads it
        from . import gamma_lut as _gamma_lut
        gamma = np.ascontiguousarray(_gamma_lut.generate(), dtype=np.uint8)
        assert gamma.shape == (8, 256)
        self._h = lib.PFCudaRendererCreate(dev, lut.ctypes.data, gamma.ctypes.data, C.byref(mode), C.byref(self._options))
        if not self._h:
            raise L.PathfinderCudaError(L.PF_CUDA_ERROR_CUDA, lib.PFCudaGetLastError().decode())

    @staticmethod
    def _make_options(dest_size, background_color):
        o = L.PFCudaRendererOptions()
        o.dest_size = L.PFVector2I(dest_size[0], dest_size[1])
        if background_color is not None:
            o.background_color = L.PFColorF(*[float(v) for v in background_color])
            o.flags = L.PF_RENDERER_OPTIONS_FLAGS_HAS_BACKGROUND_COLOR
        return o

    # -- protocol ---------------------------------------------------------------------------
    def begin_scene(self):
        L.check(L.lib().PFCudaRendererBeginScene(self._h))

    def render_command(self, command: L.PFRenderCommand):
        L.check(L.lib().PFCudaRendererRenderCommand(self._h, C.byref(command)))

    def end_scene(self):
        L.check(L.lib().PFCudaRendererEndScene(self._h))

    # -- options ----------------------------------------------------------------------------
    def set_view_box(self, view_box):
        if view_box is None:
            L.check(L.lib().PFCudaRendererSetViewBox(self._h, None))
        else:
            r = L.PFRectF(L.PFVector2F(view_box[0], view_box[1]), L.PFVector2F(view_box[2], view_box[3]))
            L.check(L.lib().PFCudaRendererSetViewBox(self._h, C.byref(r)))

    def set_strip(self, tile_y0: int, tile_y1: int):
        L.check(L.lib().PFCudaRendererSetStrip(self._h, tile_y0, tile_y1))

    # -- frame assembly across GPUs (one process per GPU; NCCL inside the library) ---------------
    def gather_init(self, gather_id: bytes, rank: int, world_size: int):
        """Collective. Joins the group and sets this renderer's strip to its share of the tile rows."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(gather_id))
        L.check(L.lib().PFCudaRendererGatherInit(self._h, buf, rank, world_size))

    GATHER_MODE_FRAME, GATHER_MODE_TILES = 0, 1

    def gather_set_mode(self, mode: int):
        """FRAME: all-gather of the finished strips; TILES: compact exports pulled over NVLink (the default when the
        GPUs can map each other's memory). Collective: every rank must choose the same mode."""
        L.check(L.lib().PFCudaRendererGatherSetMode(self._h, int(mode)))

    def gather_frame(self):
        """Collective, asynchronous: completes this rank's copy of the frame with the other ranks' strips."""
        L.check(L.lib().PFCudaRendererGatherFrame(self._h))

    def gather_wait(self):
        """Orders the renderer's stream after the gather in flight (no host wait)."""
        L.check(L.lib().PFCudaRendererGatherWait(self._h))

    def gather_destroy(self):
        L.check(L.lib().PFCudaRendererGatherDestroy(self._h))

    def set_stream(self, cuda_stream: int):
        L.check(L.lib().PFCudaRendererSetStream(self._h, cuda_stream))

    def set_dest_device_pointer(self, ptr: int, pitch: int):
        L.check(L.lib().PFCudaRendererSetDestDevicePointer(self._h, ptr, pitch))

    def dest_device_pointer(self):
        p, pitch = C.c_uint64(), C.c_size_t()
        L.check(L.lib().PFCudaRendererGetDestDevicePointer(self._h, C.byref(p), C.byref(pitch)))
        return int(p.value), int(pitch.value)

    def set_peer_dests(self, handles: list, offsets: list):
        """Registers the peers' frame buffers (IPC handles from ipc_export) for the fused gather."""
        n = len(handles)
        buf = (C.c_uint8 * (64 * max(n, 1)))()
        for i, h in enumerate(handles):
            C.memmove(C.addressof(buf) + 64 * i, bytes(h), 64)
        offs = (C.c_uint64 * max(n, 1))(*[int(o) for o in offsets])
        L.check(L.lib().PFCudaRendererSetPeerDests(self._h, buf, offs, n))

    def set_debug_lists_enabled(self, enabled: bool):
        L.check(L.lib().PFCudaRendererSetDebugListsEnabled(self._h, int(enabled)))

    def set_deferred_verification(self, enabled: bool):
        L.check(L.lib().PFCudaRendererSetDeferredVerification(self._h, int(enabled)))

    def set_timing_enabled(self, enabled: bool):
        L.check(L.lib().PFCudaRendererSetTimingEnabled(self._h, int(enabled)))

    def synchronize(self):
        L.check(L.lib().PFCudaRendererSynchronize(self._h))

    def read_texture_page(self, page_id: int) -> np.ndarray:
        """(H, W, 4) uint8 of a texture page: what a render target holds after the frame."""
        size = L.PFVector2I()
        L.check(L.lib().PFCudaRendererReadTexturePage(self._h, page_id, None, 0, C.byref(size)))
        out = np.zeros((size.y, size.x, 4), dtype=np.uint8)
        L.check(L.lib().PFCudaRendererReadTexturePage(self._h, page_id, out.ctypes.data, size.x * 4, None))
        return out

    # -- results ----------------------------------------------------------------------------
    def read_pixels(self, out: np.ndarray | None = None) -> np.ndarray:
        w, h = self.dest_size
        if out is None:
            out = np.empty((h, w, 4), dtype=np.uint8)
        L.check(L.lib().PFCudaRendererReadPixels(self._h, out.ctypes.data, w * 4))
        return out

    def read_pixels_into(self, host_ptr: int, stride: int):
        L.check(L.lib().PFCudaRendererReadPixels(self._h, host_ptr, stride))

    def stats(self) -> dict:
        s = L.PFCudaRenderStats()
        L.check(L.lib().PFCudaRendererGetStats(self._h, C.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in s._fields_}

    def times(self) -> dict:
        t = L.PFCudaRenderTime()
        L.check(L.lib().PFCudaRendererGetTimes(self._h, C.byref(t)))
        return {n: float(getattr(t, n)) for n, _ in t._fields_}

    def accumulated_times(self):
        """(stage times summed over every batch since timing was switched on, number of batches)."""
        t, n = L.PFCudaRenderTime(), C.c_uint32()
        L.check(L.lib().PFCudaRendererGetAccumulatedTimes(self._h, C.byref(t), C.byref(n)))
        return {k: float(getattr(t, k)) for k, _ in t._fields_}, int(n.value)

    # -- stage-level read-backs (parity tests) --------------------------------------------------
    @staticmethod
    def _count(n: int) -> int:
        if n < 0:
            raise L.PathfinderCudaError(int(-n), L.lib().PFCudaGetLastError().decode())
        return int(n)

    def debug_lines(self):
        lib = L.lib()
        n = self._count(lib.PFCudaRendererDebugCopyLines(self._h, None, None, 0))
        lines = np.zeros((n, 4), dtype=np.float32)
        paths = np.zeros(n, dtype=np.uint32)
        if n:
            self._count(lib.PFCudaRendererDebugCopyLines(self._h, lines.ctypes.data, paths.ctypes.data, n))
        return lines, paths

    def debug_fills(self) -> np.ndarray:
        lib = L.lib()
        n = self._count(lib.PFCudaRendererDebugCopyFills(self._h, None, 0))
        out = np.zeros(n, dtype=FILL_DTYPE)
        if n:
            self._count(lib.PFCudaRendererDebugCopyFills(self._h, out.ctypes.data, n))
        return out

    def debug_tiles(self) -> np.ndarray:
        lib = L.lib()
        n = self._count(lib.PFCudaRendererDebugCopyTiles(self._h, None, 0))
        out = np.zeros(n, dtype=TILE_DTYPE)
        if n:
            self._count(lib.PFCudaRendererDebugCopyTiles(self._h, out.ctypes.data, n))
        return out

    def debug_clips(self) -> np.ndarray:
        lib = L.lib()
        n = self._count(lib.PFCudaRendererDebugCopyClips(self._h, None, 0))
        out = np.zeros(n, dtype=CLIP_DTYPE)
        if n:
            self._count(lib.PFCudaRendererDebugCopyClips(self._h, out.ctypes.data, n))
        return out

    def debug_z_buffer(self):
        lib = L.lib()
        rect = (C.c_int32 * 4)()
        n = self._count(lib.PFCudaRendererDebugCopyZBuffer(self._h, None, 0, C.byref(rect)))
        out = np.zeros(n, dtype=np.int32)
        if n:
            self._count(lib.PFCudaRendererDebugCopyZBuffer(self._h, out.ctypes.data, n, C.byref(rect)))
        w, h = rect[2] - rect[0], rect[3] - rect[1]
        return out.reshape(h, w), tuple(rect)

    def debug_alpha_masks(self) -> np.ndarray:
        lib = L.lib()
        n = self._count(lib.PFCudaRendererDebugCopyAlphaMasks(self._h, None, 0))
        out = np.zeros((n, 16, 16), dtype=np.float32)
        if n:
            self._count(lib.PFCudaRendererDebugCopyAlphaMasks(self._h, out.ctypes.data, n))
        return out

    def close(self):
        if getattr(self, "_h", None):
            L.lib().PFCudaRendererDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
