/* include/pf_cuda.h — C ABI of the B200-native `cuda` backend for Pathfinder 3's D3D11-level
 * rasterization pipeline (bound -> dice -> bin -> propagate -> sort -> fill -> tile).
 *
 * This is the drop-in boundary: what a Rust `pathfinder_cuda` crate binds with `extern "C"`
 * (see INTEGRATION.md). Plain pointers and sizes only; no CUDA or torch types. Every entry point
 * cites the reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions follow the reference's own C API (c/src/lib.rs:564-690): opaque `*Ref` handles,
 * `Create`/`Destroy` pairs, backend-suffixed names. Unlike the reference (which panics), every
 * fallible call returns a PFCudaStatus; PFCudaGetLastError() holds the message. The Rust wrapper
 * turns a non-zero status into panic! to keep the reference's error behaviour.
 *
 * Threading: a renderer handle is single-threaded (&mut self in the reference,
 * renderer/src/gpu/renderer.rs:350-460). Command payloads are borrowed for the duration of the
 * call and copied before it returns (the reference shares them by move / Arc).
 */
#ifndef PF_CUDA_H
#define PF_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------- */
/* Plain types shared with the reference C API (c/src/lib.rs:118-175).                          */
/* ------------------------------------------------------------------------------------------- */

typedef struct PFColorF { float r, g, b, a; } PFColorF;
typedef struct PFColorU { uint8_t r, g, b, a; } PFColorU;
typedef struct PFVector2F { float x, y; } PFVector2F;
typedef struct PFVector2I { int32_t x, y; } PFVector2I;
typedef struct PFRectF { PFVector2F origin, lower_right; } PFRectF;
typedef struct PFRectI { PFVector2I origin, lower_right; } PFRectI;
/* Row-major order (c/src/lib.rs:151-163): m00 = m11(), m01 = m12(), m10 = m21(), m11 = m22(). */
typedef struct PFMatrix2x2F { float m00, m01, m10, m11; } PFMatrix2x2F;
typedef struct PFTransform2F { PFMatrix2x2F matrix; PFVector2F vector; } PFTransform2F;

typedef int32_t PFCudaStatus;
#define PF_CUDA_OK 0
#define PF_CUDA_ERROR_INVALID_ARGUMENT 1
#define PF_CUDA_ERROR_CUDA 2            /* a CUDA runtime call failed */
#define PF_CUDA_ERROR_UNSUPPORTED 3     /* command / feature outside the hot path */
#define PF_CUDA_ERROR_WRONG_LEVEL 4     /* a D3D9-only command: Renderer::require_d3d11 panics here
                                           (renderer/src/gpu/renderer.rs:1349-1360) */
#define PF_CUDA_ERROR_PROTOCOL 5        /* command outside begin_scene/end_scene, missing scene... */
#define PF_CUDA_ERROR_NO_DEVICE 6

/* Last error message of the calling thread ("" if none). Never NULL. */
const char *PFCudaGetLastError(void);

/* ------------------------------------------------------------------------------------------- */
/* #[repr(C)] records crossing the boundary (renderer/src/gpu_data.rs). Layouts are identical.  */
/* ------------------------------------------------------------------------------------------- */

#define PF_TILE_WIDTH 16   /* renderer/src/tiles.rs:19 */
#define PF_TILE_HEIGHT 16  /* renderer/src/tiles.rs:20 */
#define PF_TILE_CTRL_MASK_WINDING 0x1   /* gpu_data.rs:32 */
#define PF_TILE_CTRL_MASK_EVEN_ODD 0x2  /* gpu_data.rs:33 */
#define PF_CURVE_IS_QUADRATIC 0x80000000u /* renderer/src/builder.rs:47 */
#define PF_CURVE_IS_CUBIC 0x40000000u     /* renderer/src/builder.rs:48 */
#define PF_PATH_INDEX_NONE 0xffffffffu    /* PathBatchIndex::none(), gpu_data.rs:440-445 */
#define PF_ALPHA_TILE_ID_INVALID 0xffffffffu /* AlphaTileId::invalid(), gpu_data.rs:456-459 */

typedef struct PFSegmentIndicesD3D11 {     /* gpu_data.rs:176-181 */
    uint32_t first_point_index;
    uint32_t flags;
} PFSegmentIndicesD3D11;

typedef struct PFSegmentsD3D11 {           /* gpu_data.rs:170-174 (Vec -> ptr,len) */
    const PFVector2F *points;
    size_t point_count;
    const PFSegmentIndicesD3D11 *indices;
    size_t index_count;
} PFSegmentsD3D11;

typedef struct PFDiceMetadataD3D11 {       /* gpu_data.rs:327-334 */
    uint32_t global_path_id;
    uint32_t first_global_segment_index;
    uint32_t first_batch_segment_index;
    uint32_t pad;
} PFDiceMetadataD3D11;

typedef struct PFTilePathInfoD3D11 {       /* gpu_data.rs:297-309 */
    int16_t tile_min_x, tile_min_y, tile_max_x, tile_max_y;
    uint32_t first_tile_index;
    uint16_t color;
    uint8_t ctrl;
    int8_t backdrop;
} PFTilePathInfoD3D11;

typedef struct PFPropagateMetadataD3D11 {  /* gpu_data.rs:312-325 */
    PFRectI tile_rect;
    uint32_t tile_offset;
    uint32_t path_index;
    uint32_t z_write;
    uint32_t clip_path_index;
    uint32_t backdrop_offset;
    uint32_t pad0, pad1, pad2;
} PFPropagateMetadataD3D11;

typedef struct PFBackdropInfoD3D11 {       /* gpu_data.rs:407-414 */
    int32_t initial_backdrop;
    int32_t tile_x_offset;
    uint32_t path_index;
} PFBackdropInfoD3D11;

typedef struct PFFill {                    /* gpu_data.rs:354-363 */
    uint16_t from_x, from_y, to_x, to_y;   /* LineSegmentU16, 4.8 fixed point inside the tile */
    uint32_t link;
} PFFill;

typedef struct PFTileObjectPrimitive {     /* gpu_data.rs:264-275 */
    int16_t tile_x, tile_y;
    uint32_t alpha_tile_id;
    uint32_t path_id;
    uint16_t color;
    uint8_t ctrl;
    int8_t backdrop;
} PFTileObjectPrimitive;

typedef struct PFClip {                    /* gpu_data.rs:376-383 */
    uint32_t dest_tile_id;
    int32_t dest_backdrop;
    uint32_t src_tile_id;
    int32_t src_backdrop;
} PFClip;

/* BlendMode (content/src/effects.rs:99-163), numbered like the reference's enum. The Rust glue maps each
 * variant to these constants explicitly (it must not rely on Rust discriminants or layouts). */
#define PF_BLEND_MODE_CLEAR 0
#define PF_BLEND_MODE_COPY 1
#define PF_BLEND_MODE_SRC_IN 2
#define PF_BLEND_MODE_SRC_OUT 3
#define PF_BLEND_MODE_SRC_OVER 4
#define PF_BLEND_MODE_SRC_ATOP 5
#define PF_BLEND_MODE_DEST_IN 6
#define PF_BLEND_MODE_DEST_OUT 7
#define PF_BLEND_MODE_DEST_OVER 8
#define PF_BLEND_MODE_DEST_ATOP 9
#define PF_BLEND_MODE_XOR 10
#define PF_BLEND_MODE_LIGHTER 11
#define PF_BLEND_MODE_DARKEN 12
#define PF_BLEND_MODE_LIGHTEN 13
#define PF_BLEND_MODE_MULTIPLY 14
#define PF_BLEND_MODE_SCREEN 15
#define PF_BLEND_MODE_HARD_LIGHT 16
#define PF_BLEND_MODE_OVERLAY 17
#define PF_BLEND_MODE_COLOR_DODGE 18
#define PF_BLEND_MODE_COLOR_BURN 19
#define PF_BLEND_MODE_SOFT_LIGHT 20
#define PF_BLEND_MODE_DIFFERENCE 21
#define PF_BLEND_MODE_EXCLUSION 22
#define PF_BLEND_MODE_HUE 23
#define PF_BLEND_MODE_SATURATION 24
#define PF_BLEND_MODE_COLOR 25
#define PF_BLEND_MODE_LUMINOSITY 26

#define PF_COLOR_COMBINE_MODE_NONE 0 /* ColorCombineMode, gpu_data.rs:346-352 */
#define PF_COLOR_COMBINE_MODE_SRC_IN 1
#define PF_COLOR_COMBINE_MODE_DEST_IN 2

/* Filter (content/src/effects.rs:44-97), flattened: the Rust type is a data-carrying enum with no C layout.
 *   NONE             no parameters
 *   RADIAL_GRADIENT  params = line.from.xy, line.to.xy, radii.xy, uv_origin.xy
 *   TEXT             params = fg rgba, bg rgba, defringing kernel[4] (PF_FILTER_FLAG_TEXT_HAS_KERNEL),
 *                    gamma correction (PF_FILTER_FLAG_TEXT_GAMMA_CORRECTION)
 *   BLUR             params[0] = sigma; PF_FILTER_FLAG_BLUR_Y for BlurDirection::Y
 *   COLOR_MATRIX     params = the five F32x4 columns */
#define PF_FILTER_NONE 0
#define PF_FILTER_RADIAL_GRADIENT 1
#define PF_FILTER_TEXT 2
#define PF_FILTER_BLUR 3
#define PF_FILTER_COLOR_MATRIX 4
#define PF_FILTER_FLAG_TEXT_HAS_KERNEL 0x1
#define PF_FILTER_FLAG_TEXT_GAMMA_CORRECTION 0x2
#define PF_FILTER_FLAG_BLUR_Y 0x1
typedef struct PFFilter {
    uint32_t kind;
    uint32_t flags;
    float params[20];
} PFFilter;

/* The C-side form of TextureMetadataEntry (gpu_data.rs:336-344). NOT the memory layout of the Rust struct
 * (its Transform2F is a 16-byte aligned F32x4 + F32x2, its Filter a data-carrying enum, its BlendMode a one-byte
 * enum): the Rust glue converts every entry (integration/pathfinder_cuda/src/lib.rs
 * `texture_metadata_entry`). Evaluated: every colour combine mode, every filter, every blend mode. The destructive
 * modes (BlendMode::is_destructive: CLEAR, COPY, SRC_IN, DEST_IN, SRC_OUT, DEST_ATOP) act on the tiles the path draws —
 * its tile map covers the view box (builder.rs:430-434) but empty tiles are never drawn (builder.rs:1014-1016). */
typedef struct PFTextureMetadataEntry {
    PFTransform2F color_0_transform;
    uint32_t color_0_combine_mode;   /* PF_COLOR_COMBINE_MODE_* */
    PFColorU base_color;
    uint32_t blend_mode;             /* PF_BLEND_MODE_* */
    PFFilter filter;
} PFTextureMetadataEntry;

/* PrepareTilesInfoD3D11 (gpu_data.rs:149-168) with the Vecs flattened. `backdrops` may be NULL
 * with backdrop_count = 0, meaning "all initial backdrops are zero" (what init_backdrops,
 * renderer/src/builder.rs:762-768, produces in PrepareMode::GPU). */
typedef struct PFPrepareTilesInfoD3D11 {
    const PFBackdropInfoD3D11 *backdrops;
    size_t backdrop_count;
    const PFPropagateMetadataD3D11 *propagate_metadata;
    const PFDiceMetadataD3D11 *dice_metadata;
    const PFTilePathInfoD3D11 *tile_path_info;   /* all three: path_count entries */
    PFTransform2F transform;
} PFPrepareTilesInfoD3D11;

#define PF_PATH_SOURCE_DRAW 0  /* gpu_data.rs:142-147 */
#define PF_PATH_SOURCE_CLIP 1

typedef struct PFClippedPathInfo {         /* gpu_data.rs:183-199 */
    uint32_t clip_batch_id;
    uint32_t clipped_path_count;
    uint32_t max_clipped_tile_count;
} PFClippedPathInfo;

typedef struct PFTileBatchDataD3D11 {      /* gpu_data.rs:121-140 */
    uint32_t batch_id;
    uint32_t path_count;
    uint32_t tile_count;
    uint32_t segment_count;
    PFPrepareTilesInfoD3D11 prepare_info;
    uint32_t path_source;
    uint32_t has_clipped_path_info;
    PFClippedPathInfo clipped_path_info;
    /* Extension (not in the reference): 0 = unknown. A non-zero key equal to the previous batch's
     * promises byte-identical metadata arrays, so the renderer keeps its device copy — the
     * reference's "scene not dirty" idea (builder.rs:193-215) applied to the batch data it
     * otherwise rebuilds and re-uploads every frame. The Rust glue may always pass 0. */
    uint64_t content_key;
} PFTileBatchDataD3D11;

/* RenderCommand (gpu_data.rs:37-105) as a tagged union. */
typedef enum PFRenderCommandKind {
    PF_RENDER_COMMAND_START = 0,
    PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE = 1,
    PF_RENDER_COMMAND_UPLOAD_TEXEL_DATA = 2,
    PF_RENDER_COMMAND_DECLARE_RENDER_TARGET = 3,
    PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA = 4,
    PF_RENDER_COMMAND_ADD_FILLS_D3D9 = 5,
    PF_RENDER_COMMAND_FLUSH_FILLS_D3D9 = 6,
    PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11 = 7,
    PF_RENDER_COMMAND_PUSH_RENDER_TARGET = 8,
    PF_RENDER_COMMAND_POP_RENDER_TARGET = 9,
    PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11 = 10,
    PF_RENDER_COMMAND_DRAW_TILES_D3D9 = 11,
    PF_RENDER_COMMAND_DRAW_TILES_D3D11 = 12,
    PF_RENDER_COMMAND_FINISH = 13
} PFRenderCommandKind;

/* TextureLocation (gpu_data.rs:222-226): a rectangle of texels inside a texture page. */
typedef struct PFTextureLocation {
    uint32_t page; /* TexturePageId */
    PFRectI rect;
} PFTextureLocation;

/* TextureSamplingFlags (gpu/src/lib.rs:521-528) and TileBatchTexture (gpu_data.rs:247-254). composite_op is
 * PaintCompositeOp (paint.rs): carried, and like in the reference not read by the renderer (the texture metadata
 * entry's color_0_combine_mode says how texture and base colour combine). */
#define PF_TEXTURE_SAMPLING_FLAGS_REPEAT_U 0x1
#define PF_TEXTURE_SAMPLING_FLAGS_REPEAT_V 0x2
#define PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MIN 0x4
#define PF_TEXTURE_SAMPLING_FLAGS_NEAREST_MAG 0x8
#define PF_PAINT_COMPOSITE_OP_SRC_IN 0
#define PF_PAINT_COMPOSITE_OP_DEST_IN 1
typedef struct PFTileBatchTexture {
    uint32_t page;
    uint8_t sampling_flags;
    uint8_t composite_op;
} PFTileBatchTexture;

typedef struct PFRenderCommand {
    uint32_t kind; /* PFRenderCommandKind */
    union {
        struct { uint64_t path_count; uint32_t needs_readable_framebuffer; } start;
        /* AllocateTexturePage { page_id, descriptor } (gpu_data.rs:53-54): an RGBA8 page of `size` texels. */
        struct { uint32_t page_id; PFVector2I size; } allocate_texture_page;
        /* UploadTexelData { texels, location } (gpu_data.rs:56-57): location.rect.area() texels, row-major. */
        struct { const PFColorU *texels; size_t texel_count; PFTextureLocation location; } upload_texel_data;
        /* DeclareRenderTarget { id, location } (gpu_data.rs:59-62). Rendering to it starts from transparent black.
         * As in the reference's GL backend, texture coordinates address a render target bottom-up (v = 1 is its
         * top row: paint.rs:628-634 flips v for PatternSource::RenderTarget), pages filled by UploadTexelData
         * top-down. */
        struct { uint32_t render_target_id; PFTextureLocation location; } declare_render_target;
        struct { const PFTextureMetadataEntry *entries; size_t entry_count; uint64_t content_key; /* as
                 PFTileBatchDataD3D11.content_key */ } upload_texture_metadata;
        struct { PFSegmentsD3D11 draw_segments, clip_segments;
                 /* Extension (0 = the reference's semantics: the arrays are borrowed for the call, the
                  * renderer waits for its copies). Non-zero: the caller keeps the arrays valid and
                  * unmodified until the frame has been verified — PFCudaRendererEndScene by default,
                  * with deferred verification the next BeginScene / Synchronize / ReadPixels / GetStats
                  * on this renderer — so the host-to-device copies are enqueued without a wait. */
                 uint32_t payload_persists; } upload_scene_d3d11;
        struct { PFTileBatchDataD3D11 batch; } prepare_clip_tiles_d3d11;
        struct { PFTileBatchDataD3D11 tile_batch_data; uint32_t has_color_texture; /* Option<TileBatchTexture> */
                 PFTileBatchTexture color_texture; } draw_tiles_d3d11;
        struct { uint32_t render_target_id; } push_render_target;
        struct { uint64_t cpu_build_time_ns; } finish;
    } u;
} PFRenderCommand;

/* ------------------------------------------------------------------------------------------- */
/* Device and renderer (replaces Renderer<D: Device>, renderer/src/gpu/renderer.rs:181-460).     */
/* ------------------------------------------------------------------------------------------- */

typedef struct PFCudaDevice *PFCudaDeviceRef;
typedef struct PFCudaRenderer *PFCudaRendererRef;

#define PF_RENDERER_LEVEL_D3D9 0x1   /* c/src/lib.rs:91 */
#define PF_RENDERER_LEVEL_D3D11 0x2  /* c/src/lib.rs:92 */
typedef struct PFRendererMode { uint8_t level; } PFRendererMode; /* c/src/lib.rs:199-202 */

#define PF_RENDERER_OPTIONS_FLAGS_HAS_BACKGROUND_COLOR 0x1 /* c/src/lib.rs:88 */
#define PF_RENDERER_OPTIONS_FLAGS_SHOW_DEBUG_UI 0x2        /* c/src/lib.rs:89 (ignored) */

/* RendererOptions + DestFramebuffer::full_window (renderer/src/gpu/options.rs:19-33,79-119): the
 * destination is an RGBA8 image of dest_size pixels owned by the renderer in device memory. */
typedef struct PFCudaRendererOptions {
    PFVector2I dest_size;
    PFColorF background_color;
    uint8_t flags;
} PFCudaRendererOptions;

/* Device::feature_level() analogue (gpu/src/lib.rs:32-50): always D3D11. `ordinal` is the CUDA
 * device index (the process's LOCAL_RANK under one-process-per-GPU launches). */
PFCudaDeviceRef PFCudaDeviceCreate(int32_t ordinal);
void PFCudaDeviceDestroy(PFCudaDeviceRef device);
uint8_t PFCudaDeviceGetFeatureLevel(PFCudaDeviceRef device);

/* Renderer::new (gpu/renderer.rs:181-340). Takes ownership of `device` (as PFGLRendererCreate,
 * c/src/lib.rs:601-615). The two LUTs are the decoded `textures/area-lut.png` (256x256 RGBA8) and
 * `textures/gamma-lut.png` (256x8 L8, may be NULL: only the text filter reads it) that the
 * reference fetches through its ResourceLoader. mode->level must be D3D11. Returns NULL on error. */
PFCudaRendererRef PFCudaRendererCreate(PFCudaDeviceRef device, const uint8_t *area_lut_rgba8,
                                       const uint8_t *gamma_lut_l8, const PFRendererMode *mode,
                                       const PFCudaRendererOptions *options);
void PFCudaRendererDestroy(PFCudaRendererRef renderer);

/* Renderer::options_mut + dest_framebuffer_size_changed (gpu/renderer.rs:600-625). */
PFCudaStatus PFCudaRendererSetOptions(PFCudaRendererRef renderer, const PFCudaRendererOptions *options);

/* Renderer::begin_scene / render_command / end_scene (gpu/renderer.rs:350-460). */
PFCudaStatus PFCudaRendererBeginScene(PFCudaRendererRef renderer);
PFCudaStatus PFCudaRendererRenderCommand(PFCudaRendererRef renderer, const PFRenderCommand *command);
PFCudaStatus PFCudaRendererEndScene(PFCudaRendererRef renderer);

/* Reads the destination image back (row-major, top-left origin, `stride` bytes per row >= 4*w).
 * Synchronises the renderer's stream. Replaces Device::read_pixels (gpu/src/lib.rs:100-104). */
PFCudaStatus PFCudaRendererReadPixels(PFCudaRendererRef renderer, uint8_t *dst, size_t stride);
/* Reads a texture page back (RGBA8, row-major, rows top-down, `stride` bytes per row >= 4 * page width): what a
 * render target holds after the frame, or what UploadTexelData put there. Synchronises the renderer's stream. */
PFCudaStatus PFCudaRendererReadTexturePage(PFCudaRendererRef renderer, uint32_t page_id, uint8_t *dst, size_t stride,
                                           PFVector2I *size_out);
/* Device pointer of the destination image (for peer copies / collectives) and its row pitch. */
PFCudaStatus PFCudaRendererGetDestDevicePointer(PFCudaRendererRef renderer, uint64_t *device_ptr,
                                                size_t *pitch_bytes);
/* Redirects the destination image to caller-owned device memory (e.g. this rank's slot of an
 * all-gather buffer); pass 0 to return to the renderer-owned image. */
PFCudaStatus PFCudaRendererSetDestDevicePointer(PFCudaRendererRef renderer, uint64_t device_ptr,
                                                size_t pitch_bytes);
/* Fused all-gather for the strip partition: every rank exports its frame buffer with PFCudaIpcExport
 * (a 64-byte CUDA IPC handle + the pointer's offset inside its allocation), the handles are exchanged
 * out of band (e.g. torch.distributed.all_gather_object), and each rank registers its peers' buffers
 * here. From then on the fused fill+tile kernel stores every finished tile into the local image and
 * into all peers' images over NVLink, so no separate collective is needed — only a barrier before the
 * assembled frame is read. count = 0 removes the peers. At most 7 peers (8 GPUs). */
PFCudaStatus PFCudaIpcExport(uint64_t device_ptr, uint8_t handle_out[64], uint64_t *offset_out);
PFCudaStatus PFCudaRendererSetPeerDests(PFCudaRendererRef renderer, const uint8_t *handles,
                                        const uint64_t *offsets, int32_t count);
/* Uses the given cudaStream_t (as an integer handle) for all work; 0 = the renderer's own. */
PFCudaStatus PFCudaRendererSetStream(PFCudaRendererRef renderer, uint64_t cuda_stream);
/* In steady state (cached batch, bounds from the previous frame) the only host wait of a frame is the
 * check that no stage overflowed its bound. With deferred verification that check moves to the next
 * call on this renderer (begin_scene, ReadPixels, Synchronize, GetStats, ...), so consecutive frames
 * are enqueued without any host wait; a frame that did overflow is re-rendered then. Callers that hand
 * the destination to an external consumer (peer copy, collective) on the stream should call
 * PFCudaRendererSynchronize first if they cannot tolerate a frame being repaired late. Default: off. */
PFCudaStatus PFCudaRendererSetDeferredVerification(PFCudaRendererRef renderer, int32_t enabled);
/* Blocks until all submitted work is complete (and verified). */
PFCudaStatus PFCudaRendererSynchronize(PFCudaRendererRef renderer);

/* Multi-GPU strip partition (not in the reference): this renderer owns tile rows
 * [tile_y0, tile_y1) of the frame; everything outside is neither binned, filled nor composited.
 * tile_y0 = tile_y1 = 0 restores the full frame. */
PFCudaStatus PFCudaRendererSetStrip(PFCudaRendererRef renderer, int32_t tile_y0, int32_t tile_y1);

/* ---- Frame assembly across GPUs (SURVEY.md §8e "distributed communication backend"; new work: the reference
 * has no multi-GPU path). One process per GPU; every rank renders its strip into its own copy of the frame and the
 * copies are completed over NCCL / NVLink from inside the library, on a stream of the renderer's own that is
 * ordered after the frame's compositing — the next frame's earlier stages overlap it. NCCL is resolved at run
 * time (libnccl.so.2), so the library loads where it is absent; the calls then return PF_CUDA_ERROR_UNSUPPORTED. */
#define PF_CUDA_GATHER_ID_BYTES 128 /* ncclUniqueId */
typedef struct PFCudaGatherId {
    uint8_t bytes[PF_CUDA_GATHER_ID_BYTES];
} PFCudaGatherId;
/* The tile rows [*tile_y0, *tile_y1) rank `rank` of `world_size` owns out of `tile_rows`: as equal as whole rows allow. */
void PFCudaStripOfRank(int32_t tile_rows, int32_t rank, int32_t world_size, int32_t *tile_y0, int32_t *tile_y1);
/* On one rank; ship the bytes to the others through any host channel (ncclGetUniqueId). */
PFCudaStatus PFCudaGatherCreateId(PFCudaGatherId *id_out);
/* Collective (every rank calls it with the same id): joins the renderer to the group (ncclCommInitRank) and sets its
 * strip to PFCudaStripOfRank of the destination's tile rows. The destination must have contiguous rows (pitch = 4 * width). */
PFCudaStatus PFCudaRendererGatherInit(PFCudaRendererRef renderer, const PFCudaGatherId *id, int32_t rank,
                                      int32_t world_size);
/* How PFCudaRendererGatherFrame moves the strips (default: PF_CUDA_GATHER_MODE_TILES when the GPUs can map each
 * other's memory, else PF_CUDA_GATHER_MODE_FRAME). Collective: every rank must choose the same mode.
 *   FRAME  all-gather of the finished strips (ncclAllGather in place when the strips are equal, grouped
 *          ncclBroadcast otherwise): every rank receives the whole frame, (N - 1) / N x 4 bytes per pixel.
 *   TILES  the fill + tile kernels of every rank push a compact export of their strip — 4 bytes per single-colour tile,
 *          one contiguous 1 KB block per other tile — straight into receive slots in the other ranks' memory (CUDA IPC
 *          over NVLink) as they finish tiles; after a barrier (ncclAllReduce of one word) a kernel on each rank expands
 *          what it received into its own copy of the frame. Frames with more than one draw batch on the destination
 *          fall back to FRAME for that frame. */
#define PF_CUDA_GATHER_MODE_FRAME 0
#define PF_CUDA_GATHER_MODE_TILES 1
PFCudaStatus PFCudaRendererGatherSetMode(PFCudaRendererRef renderer, int32_t mode);
/* Collective, asynchronous: after the frame just rendered, completes every rank's copy of the frame with the other
 * ranks' strips. ReadPixels, Synchronize and the next frame's compositing wait for what they need of it. With deferred verification a frame that overflowed a
 * stage bound is repaired at the next use of the renderer, after its (stale) strip has been gathered: render with
 * verification on when every gathered frame must be final. */
PFCudaStatus PFCudaRendererGatherFrame(PFCudaRendererRef renderer);
/* Orders the renderer's stream after the gather in flight, without a host wait (for callers that continue on that
 * stream, e.g. to time the assembled frame or to hand it to another consumer). */
PFCudaStatus PFCudaRendererGatherWait(PFCudaRendererRef renderer);
/* Leaves the group (ncclCommDestroy) and restores the full-frame strip. */
PFCudaStatus PFCudaRendererGatherDestroy(PFCudaRendererRef renderer);

/* scene.view_box() as process_line_segment sees it (renderer/src/tiler.rs:194-200): segments are
 * clipped to [min_x, max_x] x (-inf, max_y]. The RenderCommand stream does not carry it; the Rust
 * glue sets it from Scene::view_box() before begin_scene. NULL = the destination rect (what the
 * reference demo uses, demo/common/src/lib.rs:910-914). */
PFCudaStatus PFCudaRendererSetViewBox(PFCudaRendererRef renderer, const PFRectF *view_box);

/* RenderStats (renderer/src/gpu/perf.rs:21-30) plus the counts of this pipeline's stages. */
typedef struct PFCudaRenderStats {
    uint64_t path_count;
    uint64_t fill_count;
    uint64_t alpha_tile_count;
    uint64_t total_tile_count;     /* dense bbox tiles (TileBatchDataD3D11.tile_count analogue) */
    uint64_t cpu_build_time_ns;
    uint64_t drawcall_count;       /* kernels launched this frame */
    uint64_t gpu_bytes_allocated;
    uint64_t gpu_bytes_committed;
    uint64_t input_segment_count;  /* SegmentIndicesD3D11 entries diced */
    uint64_t line_segment_count;   /* flattened segments entering bin ("segments") */
    uint64_t tile_list_entry_count;/* tiles surviving the z-cull, summed over framebuffer tiles */
    uint64_t column_count;
    uint64_t host_sync_count;      /* blocking read-backs this frame (1 per batch in steady state) */
    uint64_t visible_fill_count;   /* fills of tiles that survived the z-cull (read by fill+tile) */
    uint64_t h2d_bytes;            /* host-to-device bytes copied this frame */
    uint64_t batch_cache_hits;     /* batches rendered from the cached device-side metadata */
    uint64_t reruns;               /* batches re-rendered because a stage overflowed its bound */
} PFCudaRenderStats;
PFCudaStatus PFCudaRendererGetStats(PFCudaRendererRef renderer, PFCudaRenderStats *stats);

/* RenderTime (renderer/src/gpu/perf.rs:223-229) per stage, CUDA-event milliseconds of the last
 * frame. Only recorded when timing is enabled. */
typedef struct PFCudaRenderTime {
    float upload_ms, bound_ms, dice_ms, bin_ms, propagate_ms, sort_ms, fill_tile_ms, total_ms;
} PFCudaRenderTime;
PFCudaStatus PFCudaRendererSetTimingEnabled(PFCudaRendererRef renderer, int32_t enabled);
PFCudaStatus PFCudaRendererGetTimes(PFCudaRendererRef renderer, PFCudaRenderTime *times);
/* Sums over every batch since timing was switched on (per-frame times are reset by begin_scene; with deferred
 * verification a caller that never waits inside its loop reads the totals once at the end). */
PFCudaStatus PFCudaRendererGetAccumulatedTimes(PFCudaRendererRef renderer, PFCudaRenderTime *times, uint32_t *batches);

/* ------------------------------------------------------------------------------------------- */
/* Stage-level read-backs for parity tests (SURVEY.md §8b "stage-level test hooks"). They copy   */
/* the device-side lists of the LAST DrawTilesD3D11 batch in the canonical D3D9 forms the CPU     */
/* tiler emits. Each returns the element count; when `out` is non-NULL copies min(count, cap).    */
/* ------------------------------------------------------------------------------------------- */

/* Enables retention of per-stage lists (emission-ordered fills, alpha tile ids). Off by default. */
PFCudaStatus PFCudaRendererSetDebugListsEnabled(PFCudaRendererRef renderer, int32_t enabled);
/* Flattened line segments in emission order: 4 floats each + batch path index. */
int64_t PFCudaRendererDebugCopyLines(PFCudaRendererRef renderer, float *out_lines, uint32_t *out_paths,
                                     size_t cap);
/* Fills in emission order (path order, then process_line_segment order), link = alpha tile id
 * numbered in first-fill order = SequentialExecutor numbering (AddFillsD3D9 payloads). */
int64_t PFCudaRendererDebugCopyFills(PFCudaRendererRef renderer, PFFill *out, size_t cap);
/* Non-empty tiles in path order, row-major inside a path (DrawTileBatchD3D9.tiles,
 * renderer/src/builder.rs:1013-1019), path_id = global draw path id. */
int64_t PFCudaRendererDebugCopyTiles(PFCudaRendererRef renderer, PFTileObjectPrimitive *out, size_t cap);
/* Clip records of the batch (DrawTileBatchD3D9.clips, renderer/src/builder.rs:1031-1040): one per draw tile
 * whose mask is min-combined with a clip tile's mask, in tile order. With clipped paths the fill list starts
 * with the clip paths' fills and alpha tile ids count the clip tiles first (SequentialExecutor order; the clip
 * batch holds the clip paths that are used, in order of first use). */
int64_t PFCudaRendererDebugCopyClips(PFCudaRendererRef renderer, PFClip *out, size_t cap);
/* Z-buffer over the framebuffer tile rect (DrawTileBatchD3D9.z_buffer_data). rect_out = min_x,
 * min_y, max_x, max_y in tiles. */
int64_t PFCudaRendererDebugCopyZBuffer(PFCudaRendererRef renderer, int32_t *out, size_t cap,
                                       int32_t rect_out[4]);
/* Coverage masks of all alpha tiles, [alpha_tile_id][16][16] f32, unclamped, without backdrop. */
int64_t PFCudaRendererDebugCopyAlphaMasks(PFCudaRendererRef renderer, float *out, size_t cap_tiles);

/* ------------------------------------------------------------------------------------------- */
/* Scene side (renderer/src/scene.rs, options.rs). The host logic of the D3D11 level —            */
/* BuiltSegments::from_scene, TileBatchDataD3D11::push, the command order of SceneBuilder::build  */
/* (renderer/src/builder.rs:148-222,653-841,1058-1113) — mirrored in C++ because this image has    */
/* no Rust toolchain. With Rust available the reference's own Scene produces the same commands.   */
/* ------------------------------------------------------------------------------------------- */

typedef struct PFScene *PFSceneRef;
typedef struct PFBuildOptions *PFBuildOptionsRef;
typedef struct PFRenderTransform *PFRenderTransformRef;

#define PF_FILL_RULE_WINDING 0   /* content/src/fill.rs */
#define PF_FILL_RULE_EVEN_ODD 1
#define PF_POINT_FLAGS_CONTROL_POINT_0 0x1 /* content/src/outline.rs PointFlags */
#define PF_POINT_FLAGS_CONTROL_POINT_1 0x2
#define PF_CLIP_PATH_NONE 0xffffffffu

PFSceneRef PFSceneCreate(void);                 /* Scene::new, scene.rs:55-69 */
void PFSceneDestroy(PFSceneRef scene);          /* c/src/lib.rs:786 */
PFSceneRef PFSceneClone(PFSceneRef scene);      /* Scene: Clone (scene.rs:37): same content, id and epoch */
void PFSceneSetViewBox(PFSceneRef scene, const PFRectF *view_box); /* scene.rs:223-226 */
void PFSceneGetViewBox(PFSceneRef scene, PFRectF *view_box);
void PFSceneGetBounds(PFSceneRef scene, PFRectF *bounds);
/* Scene::push_paint for a solid colour (scene.rs:186-190, paint.rs Palette::push_paint dedups). */
uint16_t PFScenePushPaint(PFSceneRef scene, const PFColorU *color);
/* Scene::push_render_target(RenderTarget::new(size, name)) (scene.rs:110-119): an off-screen RGBA8 image; paths
 * pushed until the matching PFScenePopRenderTarget (scene.rs:121-123) are drawn into it. Returns the
 * RenderTargetId (PF_PATH_INDEX_NONE on a bad size). */
uint32_t PFScenePushRenderTarget(PFSceneRef scene, int32_t width, int32_t height);
void PFScenePopRenderTarget(PFSceneRef scene);
/* Scene::push_paint(&Paint::from_pattern(pattern)) for pattern = Pattern::from_render_target(id, size) with
 * pattern.apply_transform(*pattern_transform) (NULL: identity) and pattern.set_filter(filter) (NULL or kind
 * PF_FILTER_NONE: no filter; PF_FILTER_TEXT: PatternFilter::Text) — content/src/pattern.rs:105-140,
 * renderer/src/paint.rs:138-146. The pattern transform maps render-target pixels to scene coordinates. */
uint16_t PFScenePushPaintRenderTargetPattern(PFSceneRef scene, uint32_t render_target_id,
                                             const PFTransform2F *pattern_transform, const PFFilter *filter);
/* Scene::push_paint(&Paint::from_pattern(Pattern::from_image(image))) (content/src/pattern.rs:52-103,
 * renderer/src/paint.rs:138-146): `pixels` = width x height RGBA8 texels, row-major, top row first, not
 * premultiplied (copied). The pattern transform maps image pixels to scene coordinates (NULL: identity). */
#define PF_PATTERN_FLAG_REPEAT_X 0x1      /* PatternFlags, pattern.rs:40-50 */
#define PF_PATTERN_FLAG_REPEAT_Y 0x2
#define PF_PATTERN_FLAG_NO_SMOOTHING 0x4
uint16_t PFScenePushPaintImagePattern(PFSceneRef scene, const PFColorU *pixels, int32_t width, int32_t height,
                                      const PFTransform2F *pattern_transform, uint32_t flags, const PFFilter *filter);
/* Scene::push_paint(&Paint::from_gradient(gradient)) (content/src/gradient.rs:30-186, paint.rs:126-136).
 * Linear: colours run along from -> to. Radial: the circles (from, radii[0]) -> (to, radii[1]) in the space
 * `transform` maps to scene coordinates (GradientGeometry::Radial). Stops sorted by offset. */
#define PF_GRADIENT_LINEAR 0
#define PF_GRADIENT_RADIAL 1
#define PF_GRADIENT_WRAP_CLAMP 0          /* GradientWrap, gradient.rs:62-70 */
#define PF_GRADIENT_WRAP_REPEAT 1
typedef struct PFColorStop { PFColorU color; float offset; } PFColorStop; /* gradient.rs:54-60 */
typedef struct PFGradient {
    uint32_t kind, wrap;
    PFVector2F from, to;
    float radii[2];
    PFTransform2F transform;
    const PFColorStop *stops;
    size_t stop_count;
} PFGradient;
uint16_t PFScenePushPaintGradient(PFSceneRef scene, const PFGradient *gradient);
/* Scene::push_draw_path (scene.rs:77-82). The outline is given as contours of points + flags:
 * contour i owns points [contour_offsets[i], contour_offsets[i+1]). Returns the DrawPathId, or
 * PF_PATH_INDEX_NONE (see PFCudaGetLastError) and leaves the scene unchanged when the paint id or fill rule is
 * unknown or the offsets decrease — mistakes the reference panics on later, when it indexes the palette. */
uint32_t PFScenePushDrawPath(PFSceneRef scene, const PFVector2F *points, const uint8_t *point_flags,
                             const uint32_t *contour_offsets, uint32_t contour_count,
                             uint16_t paint_id, uint8_t fill_rule, uint8_t blend_mode,
                             uint32_t clip_path_id);
/* Scene::push_clip_path (scene.rs:99-106). */
uint32_t PFScenePushClipPath(PFSceneRef scene, const PFVector2F *points, const uint8_t *point_flags,
                             const uint32_t *contour_offsets, uint32_t contour_count,
                             uint8_t fill_rule, uint32_t clip_path_id);
/* Bulk form of repeated PFScenePushDrawPath calls (large synthetic scenes): path i owns contours
 * [path_contour_offsets[i], path_contour_offsets[i+1]). */
PFCudaStatus PFScenePushDrawPaths(PFSceneRef scene, const PFVector2F *points, const uint8_t *point_flags,
                                  size_t point_count, const uint32_t *contour_offsets,
                                  size_t contour_count, const uint32_t *path_contour_offsets,
                                  size_t path_count, const uint16_t *paint_ids,
                                  const uint8_t *fill_rules, const uint32_t *clip_path_ids);
uint32_t PFSceneGetDrawPathCount(PFSceneRef scene);
uint32_t PFSceneGetEpoch(PFSceneRef scene);

PFRenderTransformRef PFRenderTransformCreate2D(const PFTransform2F *transform); /* c/src/lib.rs:740 */
void PFRenderTransformDestroy(PFRenderTransformRef transform);                  /* c/src/lib.rs:752 */
PFBuildOptionsRef PFBuildOptionsCreate(void);                                   /* c/src/lib.rs:757 */
void PFBuildOptionsDestroy(PFBuildOptionsRef options);                          /* c/src/lib.rs:762 */
PFBuildOptionsRef PFBuildOptionsClone(PFBuildOptionsRef options);               /* BuildOptions: Clone (options.rs:73) */
/* Consumes the transform (c/src/lib.rs:768-772). */
void PFBuildOptionsSetTransform(PFBuildOptionsRef options, PFRenderTransformRef transform);
void PFBuildOptionsSetDilation(PFBuildOptionsRef options, const PFVector2F *dilation); /* :774 */
void PFBuildOptionsSetSubpixelAAEnabled(PFBuildOptionsRef options, int32_t enabled);   /* :780 */

/* RenderCommandListener (renderer/src/options.rs:31-49): called once per command, in the order
 * SceneBuilder::build sends them; payload pointers are valid only during the call. A non-zero
 * return aborts the build and is returned from PFSceneBuild. */
typedef PFCudaStatus (*PFRenderCommandListenerFn)(const PFRenderCommand *command, void *userdata);

/* Scene::build at RendererLevel::D3D11 with a SequentialExecutor (scene.rs:290-297). `sink_state`
 * carries SceneSink.last_scene across builds (scene.rs:384-396): pass the address of a
 * zero-initialised PFSceneSinkState kept alongside the renderer. */
typedef struct PFSceneSinkState { uint32_t has_last_scene, last_scene_id, last_scene_epoch; } PFSceneSinkState;
PFCudaStatus PFSceneBuild(PFSceneRef scene, PFBuildOptionsRef options, PFSceneSinkState *sink_state,
                          PFRenderCommandListenerFn listener, void *userdata);
/* The same for a renderer that owns the tile rows [tile_y0, tile_y1) of the frame (PFCudaRendererSetStrip, multi-GPU):
 * draw paths without a tile in those rows are left out of the uploaded segments and of the batch (their ids stay
 * global, so draw order and occlusion are those of the whole scene). PFSceneBuildAndRenderCuda applies the
 * renderer's strip by itself. Scenes with render targets, textured paints or blend modes are built whole. */
PFCudaStatus PFSceneBuildForStrip(PFSceneRef scene, PFBuildOptionsRef options, PFSceneSinkState *sink_state,
                                  PFRenderCommandListenerFn listener, void *userdata, int32_t tile_y0, int32_t tile_y1);

/* ------------------------------------------------------------------------------------------- */
/* Stroke-to-fill on the host: OutlineStrokeToFill, content/src/stroke.rs:88-448. */
/* ------------------------------------------------------------------------------------------- */

#define PF_LINE_CAP_BUTT 0    /* stroke.rs LineCap */
#define PF_LINE_CAP_SQUARE 1
#define PF_LINE_CAP_ROUND 2
#define PF_LINE_JOIN_MITER 0  /* stroke.rs LineJoin::Miter(miter_limit) */
#define PF_LINE_JOIN_BEVEL 1
#define PF_LINE_JOIN_ROUND 2

typedef struct PFStrokeStyle {            /* stroke.rs:45-53 */
    float line_width;
    uint32_t line_cap;
    uint32_t line_join;
    float miter_limit;                    /* LineJoin::Miter's ratio; the reference's default is 10 */
} PFStrokeStyle;

typedef struct PFOutline *PFOutlineRef;

/* OutlineStrokeToFill::new + offset + into_outline on an outline given as flat arrays (points, PointFlags,
 * contour_offsets[contour_count + 1], closed flag per contour). Returns NULL (see PFCudaGetLastError) on invalid
 * arguments. The stroked outline's contours are all closed; fill it with the winding rule. */
PFOutlineRef PFOutlineStrokeToFill(const PFVector2F *points, const uint8_t *point_flags,
                                   const uint32_t *contour_offsets, const uint8_t *contour_closed,
                                   uint32_t contour_count, const PFStrokeStyle *style);
uint32_t PFOutlineGetContourCount(PFOutlineRef outline);
size_t PFOutlineGetPointCount(PFOutlineRef outline);
/* points / point_flags: PFOutlineGetPointCount entries; contour_offsets: PFOutlineGetContourCount + 1. */
void PFOutlineCopy(PFOutlineRef outline, PFVector2F *points, uint8_t *point_flags, uint32_t *contour_offsets);
/* closed: one flag per contour (PFOutlineGetContourCount entries). */
void PFOutlineCopyClosed(PFOutlineRef outline, uint8_t *closed);
void PFOutlineDestroy(PFOutlineRef outline);

/* SVG path data ("M10 10 C ... z") -> outline: absolute and relative commands, implicit repetition, smooth curves,
 * arcs as cubics (one per <= 90 degrees). The reference gets this from usvg 0.9.1 (svg/src/lib.rs:386-456), which
 * is not vendored: parity unpinned, W3C SVG 1.1 rules followed. Returns NULL on malformed data. */
PFOutlineRef PFSvgPathDataToOutline(const char *path_data);

/* ------------------------------------------------------------------------------------------- */
/* Glyph outlines. The reference reads fonts through font-kit 0.6.0                  */
/* (text/src/lib.rs:80-160: Loader::glyph_for_char, advance, outline with HintingOptions::None), which */
/* is not vendored: parity unpinned upstream of the Scene. TrueType (`glyf`) outlines only.            */
/* ------------------------------------------------------------------------------------------- */

typedef struct PFFont *PFFontRef;

/* Copies `length` bytes of an sfnt file. Returns NULL (see PFCudaGetLastError) for anything but a readable
 * TrueType-flavoured font; every later read is bounds-checked against the copy. */
PFFontRef PFFontCreateFromBytes(const uint8_t *data, size_t length);
void PFFontDestroy(PFFontRef font);
uint32_t PFFontGetUnitsPerEm(PFFontRef font);                              /* Metrics::units_per_em */
uint32_t PFFontGetGlyphCount(PFFontRef font);
uint32_t PFFontGetGlyphForCodepoint(PFFontRef font, uint32_t codepoint);   /* glyph_for_char; 0 = none */
float PFFontGetGlyphAdvance(PFFontRef font, uint32_t glyph_id);            /* advance().x, font units */
/* The glyph's contours in font units, y up, as pathfinder point lists (quadratic control points flagged
 * PF_POINT_FLAGS_CONTROL_POINT_0), all closed; composite glyphs are resolved. An empty outline for a glyph
 * without contours (space); NULL for a glyph id out of range or malformed glyph data. */
PFOutlineRef PFFontGetGlyphOutline(PFFontRef font, uint32_t glyph_id);

/* Outline::dilate (content/src/outline.rs:243-249; ContourDilator, content/src/dilation.rs:34-125) in place on an
 * outline given as flat arrays: every distinct position moves along the bisector of its neighbouring edges by
 * `amount` per axis, outwards for the outline's outermost winding (Orientation::from_outline). This is the stem
 * darkening of the text path; PFSceneBuild applies it itself when the build options carry a
 * dilation, after the transform, as Scene::apply_render_options does (renderer/src/scene.rs:249-270). */
void PFOutlineDilate(PFVector2F *points, const uint32_t *contour_offsets, uint32_t contour_count,
                     const PFVector2F *amount);

/* Scene::build_and_render (scene.rs:369-378) against the CUDA renderer: begin_scene, build with a
 * listener forwarding to PFCudaRendererRenderCommand, end_scene. Borrows everything
 * (as PFSceneProxyBuildAndRenderGL, c/src/lib.rs:672-681). */
PFCudaStatus PFSceneBuildAndRenderCuda(PFSceneRef scene, PFCudaRendererRef renderer,
                                       PFBuildOptionsRef options);

/* ------------------------------------------------------------------------------------------- */
/* SceneProxy (renderer/src/concurrent/scene_proxy.rs:35-157): a scene that lives on a thread of its own. Every call
 * but the render calls returns at once; the worker thread applies them in order. One sink per proxy (scene_proxy.rs:70):
 * a proxy feeds one renderer. Commands are handed over one at a time — a command's payload points into the scene's own
 * arrays, so the worker waits in the listener while the command is consumed; what overlaps with the caller is the build
 * up to each command (flattening, batch records), not the consumption.                                                 */
/* ------------------------------------------------------------------------------------------- */
typedef struct PFSceneProxy *PFSceneProxyRef;
/* SceneProxy::from_scene (c/src/lib.rs:791-797). Consumes the scene. */
PFSceneProxyRef PFSceneProxyCreateFromScene(PFSceneRef scene);
void PFSceneProxyDestroy(PFSceneProxyRef proxy);                                   /* c/src/lib.rs:800 */
/* replace_scene (scene_proxy.rs:77-80). Consumes the new scene; the old one is destroyed by the worker. */
PFCudaStatus PFSceneProxyReplaceScene(PFSceneProxyRef proxy, PFSceneRef new_scene);
PFCudaStatus PFSceneProxySetViewBox(PFSceneProxyRef proxy, const PFRectF *view_box); /* scene_proxy.rs:83-86 */
/* build (scene_proxy.rs:89-92): queues a build; borrows the options (copied). */
PFCudaStatus PFSceneProxyBuild(PFSceneProxyRef proxy, PFBuildOptionsRef options);
/* What render does with the commands, for any consumer: the commands of the oldest queued build, in order, up to and
 * including Finish. Returns the build's status (a non-zero return of the listener aborts it). */
PFCudaStatus PFSceneProxyReceive(PFSceneProxyRef proxy, PFRenderCommandListenerFn listener, void *userdata);
/* render (scene_proxy.rs:95-105): begin_scene, the commands of the oldest queued build, end_scene. */
PFCudaStatus PFSceneProxyRenderCuda(PFSceneProxyRef proxy, PFCudaRendererRef renderer);
/* build_and_render (scene_proxy.rs:117-122; PFSceneProxyBuildAndRenderGL, c/src/lib.rs:672-681). The build is told the
 * renderer's strip (multi-GPU) and that the renderer may copy the segment arrays without waiting. */
PFCudaStatus PFSceneProxyBuildAndRenderCuda(PFSceneProxyRef proxy, PFCudaRendererRef renderer, PFBuildOptionsRef options);
/* copy_scene (scene_proxy.rs:125-130): waits for the queued calls before it, returns a clone the caller owns. */
PFSceneRef PFSceneProxyCopyScene(PFSceneProxyRef proxy);

#ifdef __cplusplus
}
#endif
#endif /* PF_CUDA_H */
