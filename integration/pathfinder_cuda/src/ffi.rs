//! Raw bindings to include/pf_cuda.h (the renderer half).
//!
//! Passed by pointer without conversion — the reference's own `#[repr(C)]` plain-old-data records, whose
//! field order, sizes and alignment pf_cuda.h mirrors one to one (checked by the `const _` size assertions
//! at the end of this file): `Vector2F` (8 bytes), `SegmentIndicesD3D11` (8), `PropagateMetadataD3D11`
//! (48), `DiceMetadataD3D11` (16), `TilePathInfoD3D11` (16), `BackdropInfoD3D11` (12), `ColorU` (4: the
//! texels of UploadTexelData).
//!
//! Converted field by field — everything else. In particular `TextureMetadataEntry` is NOT layout
//! compatible with `PFTextureMetadataEntry`: its `Transform2F` holds a 16-byte aligned `F32x4` matrix
//! (lanes m11, m21, m12, m22; geometry/src/transform2d.rs:23,134-137), its `Filter` is a data-carrying
//! enum (content/src/effects.rs:44-60) and its `BlendMode` a one-byte enum whose `SrcOver` is
//! discriminant 4, not 0 (effects.rs:99-110). lib.rs `texture_metadata_entry` builds the C record.

use pathfinder_color::ColorU;
use pathfinder_geometry::rect::RectF;
use pathfinder_geometry::vector::Vector2F;
use pathfinder_renderer::gpu_data::{BackdropInfoD3D11, DiceMetadataD3D11, PropagateMetadataD3D11};
use pathfinder_renderer::gpu_data::{SegmentIndicesD3D11, TilePathInfoD3D11};
use std::os::raw::{c_char, c_void};

pub const PF_CUDA_OK: i32 = 0;

// PFRenderCommandKind
pub const START: u32 = 0;
pub const ALLOCATE_TEXTURE_PAGE: u32 = 1;
pub const UPLOAD_TEXEL_DATA: u32 = 2;
pub const DECLARE_RENDER_TARGET: u32 = 3;
pub const UPLOAD_TEXTURE_METADATA: u32 = 4;
pub const ADD_FILLS_D3D9: u32 = 5;
pub const FLUSH_FILLS_D3D9: u32 = 6;
pub const UPLOAD_SCENE_D3D11: u32 = 7;
pub const PUSH_RENDER_TARGET: u32 = 8;
pub const POP_RENDER_TARGET: u32 = 9;
pub const PREPARE_CLIP_TILES_D3D11: u32 = 10;
pub const DRAW_TILES_D3D9: u32 = 11;
pub const DRAW_TILES_D3D11: u32 = 12;
pub const FINISH: u32 = 13;

#[repr(C)]
pub struct PFRendererMode {
    pub level: u8, // 1 = D3D9, 2 = D3D11 (uint8_t in pf_cuda.h)
}

// PF_BLEND_MODE_*, PF_COLOR_COMBINE_MODE_*, PF_FILTER_* of pf_cuda.h
pub const BLEND_MODE_CLEAR: u32 = 0;
pub const BLEND_MODE_COPY: u32 = 1;
pub const BLEND_MODE_SRC_IN: u32 = 2;
pub const BLEND_MODE_SRC_OUT: u32 = 3;
pub const BLEND_MODE_SRC_OVER: u32 = 4;
pub const BLEND_MODE_SRC_ATOP: u32 = 5;
pub const BLEND_MODE_DEST_IN: u32 = 6;
pub const BLEND_MODE_DEST_OUT: u32 = 7;
pub const BLEND_MODE_DEST_OVER: u32 = 8;
pub const BLEND_MODE_DEST_ATOP: u32 = 9;
pub const BLEND_MODE_XOR: u32 = 10;
pub const BLEND_MODE_LIGHTER: u32 = 11;
pub const BLEND_MODE_DARKEN: u32 = 12;
pub const BLEND_MODE_LIGHTEN: u32 = 13;
pub const BLEND_MODE_MULTIPLY: u32 = 14;
pub const BLEND_MODE_SCREEN: u32 = 15;
pub const BLEND_MODE_HARD_LIGHT: u32 = 16;
pub const BLEND_MODE_OVERLAY: u32 = 17;
pub const BLEND_MODE_COLOR_DODGE: u32 = 18;
pub const BLEND_MODE_COLOR_BURN: u32 = 19;
pub const BLEND_MODE_SOFT_LIGHT: u32 = 20;
pub const BLEND_MODE_DIFFERENCE: u32 = 21;
pub const BLEND_MODE_EXCLUSION: u32 = 22;
pub const BLEND_MODE_HUE: u32 = 23;
pub const BLEND_MODE_SATURATION: u32 = 24;
pub const BLEND_MODE_COLOR: u32 = 25;
pub const BLEND_MODE_LUMINOSITY: u32 = 26;
pub const COLOR_COMBINE_MODE_NONE: u32 = 0;
pub const COLOR_COMBINE_MODE_SRC_IN: u32 = 1;
pub const COLOR_COMBINE_MODE_DEST_IN: u32 = 2;
pub const FILTER_NONE: u32 = 0;
pub const FILTER_RADIAL_GRADIENT: u32 = 1;
pub const FILTER_TEXT: u32 = 2;
pub const FILTER_BLUR: u32 = 3;
pub const FILTER_COLOR_MATRIX: u32 = 4;
pub const FILTER_FLAG_TEXT_HAS_KERNEL: u32 = 0x1;
pub const FILTER_FLAG_TEXT_GAMMA_CORRECTION: u32 = 0x2;
pub const FILTER_FLAG_BLUR_Y: u32 = 0x1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFFilter {
    pub kind: u32,
    pub flags: u32,
    pub params: [f32; 20],
}

/// The C-side form of `TextureMetadataEntry` (pf_cuda.h): built field by field, never cast.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFTextureMetadataEntry {
    pub color_0_transform: [f32; 6], // m00 m01 m10 m11 tx ty
    pub color_0_combine_mode: u32,
    pub base_color: [u8; 4],
    pub blend_mode: u32,
    pub filter: PFFilter,
}

#[repr(C)]
pub struct PFCudaRendererOptions {
    pub dest_size: [i32; 2],
    pub background_color: [f32; 4],
    pub flags: u8, // bit 0: has a background colour
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFSegmentsD3D11 {
    pub points: *const Vector2F,
    pub point_count: usize,
    pub indices: *const SegmentIndicesD3D11,
    pub index_count: usize,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFPrepareTilesInfoD3D11 {
    pub backdrops: *const BackdropInfoD3D11,
    pub backdrop_count: usize,
    pub propagate_metadata: *const PropagateMetadataD3D11,
    pub dice_metadata: *const DiceMetadataD3D11,
    pub tile_path_info: *const TilePathInfoD3D11,
    pub transform: [f32; 6], // m00 m01 m10 m11 tx ty
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFClippedPathInfo {
    pub clip_batch_id: u32,
    pub clipped_path_count: u32,
    pub max_clipped_tile_count: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFTileBatchDataD3D11 {
    pub batch_id: u32,
    pub path_count: u32,
    pub tile_count: u32,
    pub segment_count: u32,
    pub prepare_info: PFPrepareTilesInfoD3D11,
    pub path_source: u32, // 0 draw, 1 clip
    pub has_clipped_path_info: u32,
    pub clipped_path_info: PFClippedPathInfo,
    pub content_key: u64, // 0: no promise about the batch being unchanged
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFStart {
    pub path_count: u64,
    pub needs_readable_framebuffer: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFUploadTextureMetadata {
    pub entries: *const PFTextureMetadataEntry,
    pub entry_count: usize,
    pub content_key: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFUploadSceneD3D11 {
    pub draw_segments: PFSegmentsD3D11,
    pub clip_segments: PFSegmentsD3D11,
    pub payload_persists: u32, // 0: borrowed for the call (the reference's semantics)
}

// TextureLocation / TileBatchTexture / the texture-page and render-target commands (pf_cuda.h). The reference's
// RectI is an I32x4 (16-byte aligned SIMD lanes: origin x, y, lower right x, y), so locations are converted too.
pub const TEXTURE_SAMPLING_FLAGS_REPEAT_U: u8 = 0x1;
pub const TEXTURE_SAMPLING_FLAGS_REPEAT_V: u8 = 0x2;
pub const TEXTURE_SAMPLING_FLAGS_NEAREST_MIN: u8 = 0x4;
pub const TEXTURE_SAMPLING_FLAGS_NEAREST_MAG: u8 = 0x8;
pub const PAINT_COMPOSITE_OP_SRC_IN: u8 = 0;
pub const PAINT_COMPOSITE_OP_DEST_IN: u8 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFTextureLocation {
    pub page: u32,
    pub rect: [i32; 4], // origin x, y, lower right x, y
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFTileBatchTexture {
    pub page: u32,
    pub sampling_flags: u8,
    pub composite_op: u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFAllocateTexturePage {
    pub page_id: u32,
    pub size: [i32; 2],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFUploadTexelData {
    pub texels: *const ColorU, // four bytes r, g, b, a: PFColorU
    pub texel_count: usize,
    pub location: PFTextureLocation,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFDeclareRenderTarget {
    pub render_target_id: u32,
    pub location: PFTextureLocation,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PFDrawTilesD3D11 {
    pub tile_batch_data: PFTileBatchDataD3D11,
    pub has_color_texture: u32,
    pub color_texture: PFTileBatchTexture,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub union PFRenderCommandPayload {
    pub start: PFStart,
    pub allocate_texture_page: PFAllocateTexturePage,
    pub upload_texel_data: PFUploadTexelData,
    pub declare_render_target: PFDeclareRenderTarget,
    pub upload_texture_metadata: PFUploadTextureMetadata,
    pub upload_scene_d3d11: PFUploadSceneD3D11,
    pub prepare_clip_tiles_d3d11: PFTileBatchDataD3D11,
    pub draw_tiles_d3d11: PFDrawTilesD3D11,
    pub push_render_target: u32,
    pub finish_cpu_build_time_ns: u64,
}

#[repr(C)]
pub struct PFRenderCommand {
    pub kind: u32,
    pub u: PFRenderCommandPayload,
}

extern "C" {
    pub fn PFCudaGetLastError() -> *const c_char;
    pub fn PFCudaDeviceCreate(ordinal: i32) -> *mut c_void;
    pub fn PFCudaRendererCreate(device: *mut c_void, area_lut_rgba8: *const u8, gamma_lut_l8: *const u8,
                                mode: *const PFRendererMode, options: *const PFCudaRendererOptions)
                                -> *mut c_void;
    pub fn PFCudaRendererDestroy(renderer: *mut c_void);
    pub fn PFCudaRendererSetViewBox(renderer: *mut c_void, view_box: *const RectF) -> i32;
    pub fn PFCudaRendererSetDeferredVerification(renderer: *mut c_void, enabled: i32) -> i32;
    pub fn PFCudaRendererBeginScene(renderer: *mut c_void) -> i32;
    pub fn PFCudaRendererRenderCommand(renderer: *mut c_void, command: *const PFRenderCommand) -> i32;
    pub fn PFCudaRendererEndScene(renderer: *mut c_void) -> i32;
    pub fn PFCudaRendererReadPixels(renderer: *mut c_void, dst: *mut u8, stride: usize) -> i32;
    pub fn PFCudaRendererSynchronize(renderer: *mut c_void) -> i32;
    pub fn PFCudaRendererReadTexturePage(renderer: *mut c_void, page_id: u32, dst: *mut u8, stride: usize,
                                         size_out: *mut i32) -> i32; // size_out: PFVector2I, two i32
    // Several GPUs, one frame (pf_cuda.h "Frame assembly"): rank 0 makes the id, the application ships its 128 bytes.
    pub fn PFCudaGatherCreateId(id_out: *mut u8) -> i32; // PFCudaGatherId: 128 opaque bytes
    pub fn PFCudaStripOfRank(tile_rows: i32, rank: i32, world: i32, y0: *mut i32, y1: *mut i32);
    pub fn PFCudaRendererGatherInit(renderer: *mut c_void, id: *const u8, rank: i32, world: i32) -> i32;
    pub fn PFCudaRendererGatherSetMode(renderer: *mut c_void, mode: i32) -> i32;
    pub fn PFCudaRendererGatherFrame(renderer: *mut c_void) -> i32;
    pub fn PFCudaRendererGatherWait(renderer: *mut c_void) -> i32;
    pub fn PFCudaRendererGatherDestroy(renderer: *mut c_void) -> i32;
}

// Layout checks (compile time): the C records this file declares, and the reference records it passes by
// pointer, have the sizes pf_cuda.h states. A mismatch fails the build instead of rendering garbage.
const _: () = assert!(std::mem::size_of::<PFRendererMode>() == 1);
const _: () = assert!(std::mem::size_of::<PFFilter>() == 88);
const _: () = assert!(std::mem::size_of::<PFTextureMetadataEntry>() == 124);
const _: () = assert!(std::mem::size_of::<PFTextureLocation>() == 20);
const _: () = assert!(std::mem::size_of::<PFTileBatchTexture>() == 8);
const _: () = assert!(std::mem::size_of::<ColorU>() == 4);
const _: () = assert!(std::mem::size_of::<Vector2F>() == 8);
const _: () = assert!(std::mem::size_of::<SegmentIndicesD3D11>() == 8);
const _: () = assert!(std::mem::size_of::<PropagateMetadataD3D11>() == 48);
const _: () = assert!(std::mem::size_of::<DiceMetadataD3D11>() == 16);
const _: () = assert!(std::mem::size_of::<TilePathInfoD3D11>() == 16);
const _: () = assert!(std::mem::size_of::<BackdropInfoD3D11>() == 12);
