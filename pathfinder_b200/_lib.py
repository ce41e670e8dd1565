"""ctypes bindings of the C ABI in include/pf_cuda.h (libpf_cuda.so, built in-tree by
pathfinder_b200/csrc/Makefile). There is no CPU fallback: importing this module fails loudly when
the CUDA library is missing, and creating a device fails when no GPU is present."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PF_CUDA_LIB selects another build of the same library (A/B measurements); default: the in-tree build.
LIB_PATH = os.environ.get("PF_CUDA_LIB") or os.path.join(_HERE, "libpf_cuda.so")


class PFColorF(C.Structure):
    _fields_ = [("r", C.c_float), ("g", C.c_float), ("b", C.c_float), ("a", C.c_float)]


class PFColorU(C.Structure):
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("a", C.c_uint8)]


class PFVector2F(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class PFVector2I(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32)]


class PFRectF(C.Structure):
    _fields_ = [("origin", PFVector2F), ("lower_right", PFVector2F)]


class PFRectI(C.Structure):
    _fields_ = [("origin", PFVector2I), ("lower_right", PFVector2I)]


class PFMatrix2x2F(C.Structure):
    _fields_ = [("m00", C.c_float), ("m01", C.c_float), ("m10", C.c_float), ("m11", C.c_float)]


class PFTransform2F(C.Structure):
    _fields_ = [("matrix", PFMatrix2x2F), ("vector", PFVector2F)]


class PFSegmentIndicesD3D11(C.Structure):
    _fields_ = [("first_point_index", C.c_uint32), ("flags", C.c_uint32)]


class PFSegmentsD3D11(C.Structure):
    _fields_ = [("points", C.c_void_p), ("point_count", C.c_size_t), ("indices", C.c_void_p),
                ("index_count", C.c_size_t)]


class PFDiceMetadataD3D11(C.Structure):
    _fields_ = [("global_path_id", C.c_uint32), ("first_global_segment_index", C.c_uint32),
                ("first_batch_segment_index", C.c_uint32), ("pad", C.c_uint32)]


class PFTilePathInfoD3D11(C.Structure):
    _fields_ = [("tile_min_x", C.c_int16), ("tile_min_y", C.c_int16), ("tile_max_x", C.c_int16),
                ("tile_max_y", C.c_int16), ("first_tile_index", C.c_uint32), ("color", C.c_uint16),
                ("ctrl", C.c_uint8), ("backdrop", C.c_int8)]


class PFPropagateMetadataD3D11(C.Structure):
    _fields_ = [("tile_rect", PFRectI), ("tile_offset", C.c_uint32), ("path_index", C.c_uint32),
                ("z_write", C.c_uint32), ("clip_path_index", C.c_uint32),
                ("backdrop_offset", C.c_uint32), ("pad0", C.c_uint32), ("pad1", C.c_uint32),
                ("pad2", C.c_uint32)]


class PFBackdropInfoD3D11(C.Structure):
    _fields_ = [("initial_backdrop", C.c_int32), ("tile_x_offset", C.c_int32),
                ("path_index", C.c_uint32)]


PF_COLOR_COMBINE_MODE_NONE, PF_COLOR_COMBINE_MODE_SRC_IN, PF_COLOR_COMBINE_MODE_DEST_IN = 0, 1, 2
PF_FILTER_NONE, PF_FILTER_RADIAL_GRADIENT, PF_FILTER_TEXT, PF_FILTER_BLUR, PF_FILTER_COLOR_MATRIX = 0, 1, 2, 3, 4
PF_FILTER_FLAG_BLUR_Y = 0x1
PF_PATTERN_FLAG_REPEAT_X, PF_PATTERN_FLAG_REPEAT_Y, PF_PATTERN_FLAG_NO_SMOOTHING = 0x1, 0x2, 0x4
PF_GRADIENT_LINEAR, PF_GRADIENT_RADIAL = 0, 1
PF_GRADIENT_WRAP_CLAMP, PF_GRADIENT_WRAP_REPEAT = 0, 1
PF_TEXTURE_SAMPLING_FLAGS_REPEAT_U, PF_TEXTURE_SAMPLING_FLAGS_REPEAT_V = 0x1, 0x2
PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MIN, PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MAG = 0x4, 0x8
PF_FILTER_FLAG_TEXT_HAS_KERNEL, PF_FILTER_FLAG_TEXT_GAMMA_CORRECTION = 0x1, 0x2


class PFFilter(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("flags", C.c_uint32), ("params", C.c_float * 20)]


class PFColorStop(C.Structure):
    _fields_ = [("color", PFColorU), ("offset", C.c_float)]


class PFGradient(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("wrap", C.c_uint32), ("from_", PFVector2F), ("to", PFVector2F),
                ("radii", C.c_float * 2), ("transform", PFTransform2F), ("stops", C.POINTER(PFColorStop)),
                ("stop_count", C.c_size_t)]


class PFTextureMetadataEntry(C.Structure):
    _fields_ = [("color_0_transform", PFTransform2F), ("color_0_combine_mode", C.c_uint32),
                ("base_color", PFColorU), ("blend_mode", C.c_uint32), ("filter", PFFilter)]


class PFPrepareTilesInfoD3D11(C.Structure):
    _fields_ = [("backdrops", C.c_void_p), ("backdrop_count", C.c_size_t),
                ("propagate_metadata", C.c_void_p), ("dice_metadata", C.c_void_p),
                ("tile_path_info", C.c_void_p), ("transform", PFTransform2F)]


class PFStrokeStyle(C.Structure):
    _fields_ = [("line_width", C.c_float), ("line_cap", C.c_uint32), ("line_join", C.c_uint32), ("miter_limit", C.c_float)]


class PFClippedPathInfo(C.Structure):
    _fields_ = [("clip_batch_id", C.c_uint32), ("clipped_path_count", C.c_uint32),
                ("max_clipped_tile_count", C.c_uint32)]


class PFTileBatchDataD3D11(C.Structure):
    _fields_ = [("batch_id", C.c_uint32), ("path_count", C.c_uint32), ("tile_count", C.c_uint32),
                ("segment_count", C.c_uint32), ("prepare_info", PFPrepareTilesInfoD3D11),
                ("path_source", C.c_uint32), ("has_clipped_path_info", C.c_uint32),
                ("clipped_path_info", PFClippedPathInfo), ("content_key", C.c_uint64)]


class _Start(C.Structure):
    _fields_ = [("path_count", C.c_uint64), ("needs_readable_framebuffer", C.c_uint32)]


class _UploadTextureMetadata(C.Structure):
    _fields_ = [("entries", C.c_void_p), ("entry_count", C.c_size_t), ("content_key", C.c_uint64)]


class _UploadSceneD3D11(C.Structure):
    _fields_ = [("draw_segments", PFSegmentsD3D11), ("clip_segments", PFSegmentsD3D11),
                ("payload_persists", C.c_uint32)]


class _PrepareClipTiles(C.Structure):
    _fields_ = [("batch", PFTileBatchDataD3D11)]


class PFTextureLocation(C.Structure):
    _fields_ = [("page", C.c_uint32), ("rect", PFRectI)]


class PFTileBatchTexture(C.Structure):
    _fields_ = [("page", C.c_uint32), ("sampling_flags", C.c_uint8), ("composite_op", C.c_uint8)]


class _AllocateTexturePage(C.Structure):
    _fields_ = [("page_id", C.c_uint32), ("size", PFVector2I)]


class _UploadTexelData(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("texel_count", C.c_size_t), ("location", PFTextureLocation)]


class _DeclareRenderTarget(C.Structure):
    _fields_ = [("render_target_id", C.c_uint32), ("location", PFTextureLocation)]


class _DrawTilesD3D11(C.Structure):
    _fields_ = [("tile_batch_data", PFTileBatchDataD3D11), ("has_color_texture", C.c_uint32),
                ("color_texture", PFTileBatchTexture)]


class _PushRenderTarget(C.Structure):
    _fields_ = [("render_target_id", C.c_uint32)]


class _Finish(C.Structure):
    _fields_ = [("cpu_build_time_ns", C.c_uint64)]


class _CommandUnion(C.Union):
    _fields_ = [("start", _Start), ("allocate_texture_page", _AllocateTexturePage),
                ("upload_texel_data", _UploadTexelData), ("declare_render_target", _DeclareRenderTarget),
                ("upload_texture_metadata", _UploadTextureMetadata),
                ("upload_scene_d3d11", _UploadSceneD3D11),
                ("prepare_clip_tiles_d3d11", _PrepareClipTiles),
                ("draw_tiles_d3d11", _DrawTilesD3D11), ("push_render_target", _PushRenderTarget),
                ("finish", _Finish)]


class PFRenderCommand(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("u", _CommandUnion)]


(PF_RENDER_COMMAND_START, PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE, PF_RENDER_COMMAND_UPLOAD_TEXEL_DATA,
 PF_RENDER_COMMAND_DECLARE_RENDER_TARGET, PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA,
 PF_RENDER_COMMAND_ADD_FILLS_D3D9, PF_RENDER_COMMAND_FLUSH_FILLS_D3D9,
 PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11, PF_RENDER_COMMAND_PUSH_RENDER_TARGET,
 PF_RENDER_COMMAND_POP_RENDER_TARGET, PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11,
 PF_RENDER_COMMAND_DRAW_TILES_D3D9, PF_RENDER_COMMAND_DRAW_TILES_D3D11,
 PF_RENDER_COMMAND_FINISH) = range(14)

COMMAND_NAMES = ["Start", "AllocateTexturePage", "UploadTexelData", "DeclareRenderTarget",
                 "UploadTextureMetadata", "AddFillsD3D9", "FlushFillsD3D9", "UploadSceneD3D11",
                 "PushRenderTarget", "PopRenderTarget", "PrepareClipTilesD3D11", "DrawTilesD3D9",
                 "DrawTilesD3D11", "Finish"]

PF_CUDA_OK = 0
PF_CUDA_ERROR_INVALID_ARGUMENT = 1
PF_CUDA_ERROR_CUDA = 2
PF_CUDA_ERROR_UNSUPPORTED = 3
PF_CUDA_ERROR_WRONG_LEVEL = 4
PF_CUDA_ERROR_PROTOCOL = 5
PF_CUDA_ERROR_NO_DEVICE = 6

PF_RENDERER_LEVEL_D3D9 = 1
PF_RENDERER_LEVEL_D3D11 = 2
PF_RENDERER_OPTIONS_FLAGS_HAS_BACKGROUND_COLOR = 1


class PFRendererMode(C.Structure):
    _fields_ = [("level", C.c_uint8)]


class PFCudaRendererOptions(C.Structure):
    _fields_ = [("dest_size", PFVector2I), ("background_color", PFColorF), ("flags", C.c_uint8)]


class PFCudaRenderStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "path_count", "fill_count", "alpha_tile_count", "total_tile_count", "cpu_build_time_ns",
        "drawcall_count", "gpu_bytes_allocated", "gpu_bytes_committed", "input_segment_count",
        "line_segment_count", "tile_list_entry_count", "column_count", "host_sync_count",
        "visible_fill_count", "h2d_bytes", "batch_cache_hits", "reruns")]


class PFCudaRenderTime(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("upload_ms", "bound_ms", "dice_ms", "bin_ms", "propagate_ms",
                                         "sort_ms", "fill_tile_ms", "total_ms")]


class PFSceneSinkState(C.Structure):
    _fields_ = [("has_last_scene", C.c_uint32), ("last_scene_id", C.c_uint32),
                ("last_scene_epoch", C.c_uint32)]


LISTENER_FN = C.CFUNCTYPE(C.c_int32, C.POINTER(PFRenderCommand), C.c_void_p)

# name -> (restype, argtypes); every symbol include/pf_cuda.h declares.
SIGNATURES = {
    "PFCudaGetLastError": (C.c_char_p, []),
    "PFCudaDeviceCreate": (C.c_void_p, [C.c_int32]),
    "PFCudaDeviceDestroy": (None, [C.c_void_p]),
    "PFCudaDeviceGetFeatureLevel": (C.c_uint8, [C.c_void_p]),
    "PFCudaRendererCreate": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PFRendererMode),
                                          C.POINTER(PFCudaRendererOptions)]),
    "PFCudaRendererDestroy": (None, [C.c_void_p]),
    "PFCudaRendererSetOptions": (C.c_int32, [C.c_void_p, C.POINTER(PFCudaRendererOptions)]),
    "PFCudaRendererBeginScene": (C.c_int32, [C.c_void_p]),
    "PFCudaRendererRenderCommand": (C.c_int32, [C.c_void_p, C.POINTER(PFRenderCommand)]),
    "PFCudaRendererEndScene": (C.c_int32, [C.c_void_p]),
    "PFCudaRendererReadPixels": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "PFCudaRendererGetDestDevicePointer": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)]),
    "PFCudaRendererSetDestDevicePointer": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_size_t]),
    "PFCudaIpcExport": (C.c_int32, [C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]),
    "PFCudaRendererSetPeerDests": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "PFCudaRendererSetStream": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "PFCudaRendererSynchronize": (C.c_int32, [C.c_void_p]),
    "PFCudaRendererSetDeferredVerification": (C.c_int32, [C.c_void_p, C.c_int32]),
    "PFCudaRendererReadTexturePage": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(PFVector2I)]),
    "PFCudaRendererSetStrip": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    "PFCudaStripOfRank": (None, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "PFCudaGatherCreateId": (C.c_int32, [C.c_void_p]),
    "PFCudaRendererGatherInit": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "PFCudaRendererGatherSetMode": (C.c_int32, [C.c_void_p, C.c_int32]),
    "PFCudaRendererGatherFrame": (C.c_int32, [C.c_void_p]),
    "PFCudaRendererGatherWait": (C.c_int32, [C.c_void_p]),
    "PFCudaRendererGatherDestroy": (C.c_int32, [C.c_void_p]),
    "PFCudaRendererSetViewBox": (C.c_int32, [C.c_void_p, C.POINTER(PFRectF)]),
    "PFCudaRendererGetStats": (C.c_int32, [C.c_void_p, C.POINTER(PFCudaRenderStats)]),
    "PFCudaRendererSetTimingEnabled": (C.c_int32, [C.c_void_p, C.c_int32]),
    "PFCudaRendererGetTimes": (C.c_int32, [C.c_void_p, C.POINTER(PFCudaRenderTime)]),
    "PFCudaRendererGetAccumulatedTimes": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "PFCudaRendererSetDebugListsEnabled": (C.c_int32, [C.c_void_p, C.c_int32]),
    "PFCudaRendererDebugCopyLines": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "PFCudaRendererDebugCopyFills": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "PFCudaRendererDebugCopyTiles": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "PFCudaRendererDebugCopyClips": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "PFCudaRendererDebugCopyZBuffer": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32 * 4)]),
    "PFCudaRendererDebugCopyAlphaMasks": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "PFSceneCreate": (C.c_void_p, []),
    "PFSceneDestroy": (None, [C.c_void_p]),
    "PFSceneSetViewBox": (None, [C.c_void_p, C.POINTER(PFRectF)]),
    "PFSceneGetViewBox": (None, [C.c_void_p, C.POINTER(PFRectF)]),
    "PFSceneGetBounds": (None, [C.c_void_p, C.POINTER(PFRectF)]),
    "PFScenePushPaint": (C.c_uint16, [C.c_void_p, C.POINTER(PFColorU)]),
    "PFScenePushRenderTarget": (C.c_uint32, [C.c_void_p, C.c_int32, C.c_int32]),
    "PFScenePopRenderTarget": (None, [C.c_void_p]),
    "PFScenePushPaintRenderTargetPattern": (C.c_uint16, [C.c_void_p, C.c_uint32, C.POINTER(PFTransform2F), C.POINTER(PFFilter)]),
    "PFScenePushPaintImagePattern": (C.c_uint16, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(PFTransform2F),
                                                  C.c_uint32, C.POINTER(PFFilter)]),
    "PFScenePushPaintGradient": (C.c_uint16, [C.c_void_p, C.POINTER(PFGradient)]),
    "PFScenePushDrawPath": (C.c_uint32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                         C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint32]),
    "PFOutlineStrokeToFill": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "PFOutlineGetContourCount": (C.c_uint32, [C.c_void_p]),
    "PFOutlineGetPointCount": (C.c_size_t, [C.c_void_p]),
    "PFOutlineCopy": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "PFOutlineCopyClosed": (None, [C.c_void_p, C.c_void_p]),
    "PFOutlineDestroy": (None, [C.c_void_p]),
    "PFSvgPathDataToOutline": (C.c_void_p, [C.c_char_p]),
    "PFOutlineDilate": (None, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "PFFontCreateFromBytes": (C.c_void_p, [C.c_char_p, C.c_size_t]),
    "PFFontDestroy": (None, [C.c_void_p]),
    "PFFontGetUnitsPerEm": (C.c_uint32, [C.c_void_p]),
    "PFFontGetGlyphCount": (C.c_uint32, [C.c_void_p]),
    "PFFontGetGlyphForCodepoint": (C.c_uint32, [C.c_void_p, C.c_uint32]),
    "PFFontGetGlyphAdvance": (C.c_float, [C.c_void_p, C.c_uint32]),
    "PFFontGetGlyphOutline": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "PFScenePushClipPath": (C.c_uint32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                         C.c_uint8, C.c_uint32]),
    "PFScenePushDrawPaths": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "PFSceneGetDrawPathCount": (C.c_uint32, [C.c_void_p]),
    "PFSceneGetEpoch": (C.c_uint32, [C.c_void_p]),
    "PFRenderTransformCreate2D": (C.c_void_p, [C.POINTER(PFTransform2F)]),
    "PFRenderTransformDestroy": (None, [C.c_void_p]),
    "PFBuildOptionsCreate": (C.c_void_p, []),
    "PFBuildOptionsDestroy": (None, [C.c_void_p]),
    "PFBuildOptionsSetTransform": (None, [C.c_void_p, C.c_void_p]),
    "PFBuildOptionsSetDilation": (None, [C.c_void_p, C.POINTER(PFVector2F)]),
    "PFBuildOptionsSetSubpixelAAEnabled": (None, [C.c_void_p, C.c_int32]),
    "PFSceneBuild": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(PFSceneSinkState), LISTENER_FN, C.c_void_p]),
    "PFSceneBuildForStrip": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(PFSceneSinkState), LISTENER_FN, C.c_void_p,
                                         C.c_int32, C.c_int32]),
    "PFSceneBuildAndRenderCuda": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "PFSceneClone": (C.c_void_p, [C.c_void_p]),
    "PFBuildOptionsClone": (C.c_void_p, [C.c_void_p]),
    "PFSceneProxyCreateFromScene": (C.c_void_p, [C.c_void_p]),
    "PFSceneProxyDestroy": (None, [C.c_void_p]),
    "PFSceneProxyReplaceScene": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "PFSceneProxySetViewBox": (C.c_int32, [C.c_void_p, C.POINTER(PFRectF)]),
    "PFSceneProxyBuild": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "PFSceneProxyReceive": (C.c_int32, [C.c_void_p, LISTENER_FN, C.c_void_p]),
    "PFSceneProxyRenderCuda": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "PFSceneProxyBuildAndRenderCuda": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "PFSceneProxyCopyScene": (C.c_void_p, [C.c_void_p]),
}

_lib = None


class PathfinderCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[status {status}] {message}")
        self.status = status


def lib():
    """Loads libpf_cuda.so. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C pathfinder_b200/csrc`. There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != PF_CUDA_OK:
        raise PathfinderCudaError(status, lib().PFCudaGetLastError().decode("utf-8", "replace"))
