// pathfinder_b200/csrc/kernels.cuh — device-side records and launch wrappers of the pipeline
// stages (bound -> dice -> bin -> propagate -> sort -> fill+tile).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pf {

// Per-path record of a batch, built on the host from PropagateMetadataD3D11 + DiceMetadataD3D11 +
// TilePathInfoD3D11 (renderer/src/gpu_data.rs:297-334) with device-side offsets. 48 bytes,
// 16-byte aligned so a thread loads it as three 128-bit words.
struct __align__(16) PathInfo {
    int32_t min_x, min_y, max_x, max_y; // tile rect, already restricted to this renderer's strip
    uint32_t tile_offset;               // first dense tile of the path
    uint32_t col_offset;                // first column backdrop of the path
    uint32_t seg_batch_first;           // DiceMetadataD3D11.first_batch_segment_index
    uint32_t seg_global_first;          // DiceMetadataD3D11.first_global_segment_index
    uint32_t global_path_id;            // DiceMetadataD3D11.global_path_id (draw path id)
    uint32_t paint_ctrl;                // color u16 | ctrl u8 << 16 | z_write << 24 | PATH_TEXTURED
    uint32_t clip_path_index;           // PropagateMetadataD3D11.clip_path_index
    uint32_t pad;
};
static_assert(sizeof(PathInfo) == 48, "PathInfo layout");

constexpr uint32_t PATH_TEXTURED = 1u << 25; // PathInfo.paint_ctrl: the paint samples a colour texture (per-pixel colour)

struct Transform {
    float m11, m21, m12, m22, tx, ty;
    int identity;
};

struct ViewBox {
    float min_x, min_y, max_x, max_y;
};

struct FbRect {
    int32_t min_x, min_y, max_x, max_y; // framebuffer tile rect = round_out(view_box / 16)
};

// One entry of a framebuffer tile's painter's-order list (32 bytes, two 128-bit loads). The paint's
// colour is stored inline so that the fused kernel has no dependent entry -> paint-table load.
struct __align__(32) TileEntry {
    uint32_t fill_end;    // end of the tile's run in the tile-grouped fill array
    uint32_t word;        // fill count (low 24 bits) | backdrop i8 << 24
    uint32_t paint_ctrl;  // color u16 | ctrl u8 << 16
    uint32_t tile_index;  // dense tile index: ascending = draw order (sort key inside a list)
    float4 color;         // the paint premultiplied, (rgb * a, a); base colour already rounded through f16
};

// Tile-grouped fill: the 4.8 fixed point segment (LineSegmentU16) as one 64-bit word.
typedef uint2 PackedFill; // x = from_x | from_y << 16, y = to_x | to_y << 16

struct EmitFill { // emission-ordered fill kept for parity dumps (12 bytes like gpu_data.rs Fill)
    uint32_t from, to, tile;
};

// Coarse index over one of the per-path offset arrays: table[k] = largest path p with
// offsets[p] <= (k << shift); (count >> shift) + 2 entries. Built on the host with the batch.
struct CoarseIndex {
    const uint32_t *table;
    int shift;
};

// Everything the stage kernels read about one batch. Search arrays have n_paths + 1 entries (the
// last one is the total) so a thread finds its path with one binary search over a dense array.
struct BatchDev {
    const float2 *points;
    const uint2 *seg_indices;
    const PathInfo *paths;
    const uint32_t *path_seg_first;
    const uint32_t *path_tile_offset;
    const uint32_t *path_col_offset;
    CoarseIndex seg_index, tile_index, col_index;
    uint32_t n_paths, n_segments, n_tiles, n_columns;
    Transform xf;
    ViewBox view_box;
    FbRect fb;
};

// ---- stage launchers (all asynchronous on `stream`; return the number of kernels launched) ----

// dice: count pass writes seg_line_count[s]; emit pass reads seg_line_offset[s] and writes
// lines/line_path.
int launch_dice(bool emit, const BatchDev &b, uint32_t *seg_line_count, const uint32_t *seg_line_offset,
                float4 *lines, uint32_t *line_path, uint32_t line_capacity, cudaStream_t stream);
// Steady state: one pass, lines appended in arbitrary order; *line_count (zeroed by the caller) ends as the total.
int launch_dice_stream(const BatchDev &b, float4 *lines, uint32_t *line_path, uint32_t line_capacity,
                       uint32_t *line_count, cudaStream_t stream);

struct BinArgs {
    const float4 *lines;
    const uint32_t *line_path;
    uint32_t n_lines;            // host-side bound (grid size)
    const uint32_t *n_lines_dev; // optional: actual count on the device, min(*n_lines_dev, n_lines) is used
    uint32_t *tile_word;         // count (24) | backdrop delta (8)
    int32_t *col_backdrop;
    // BIN_COUNT: optional fills per line (parity dumps), for emission-order offsets
    uint32_t *line_fill_count;
    // BIN_EMIT_LIVE / BIN_EMIT
    const uint32_t *tile_fb;     // BIN_EMIT_LIVE: 0xffffffff = tile culled, its fills are not stored
    const uint32_t *path_live;   // BIN_EMIT_LIVE: 0 = every tile of the path was culled, skip its lines
    uint32_t *tile_fill_pos;     // running cursor, initialised with the exclusive scan of the (live) counts
    PackedFill *fills;           // tile-grouped
    uint32_t fill_capacity;
    // BIN_EMIT only (parity dumps)
    const uint32_t *line_fill_offset;
    uint32_t *tile_first_fill;   // min emission index per tile
    EmitFill *fills_emit;        // every fill in emission order
    uint32_t emit_capacity;
    // Lines crossing many tiles are queued here by k_bin and walked by whole warps in k_bin_long.
    uint32_t *long_queue;        // line indices (NULL: walk everything in k_bin)
    uint32_t long_capacity;
    uint32_t *long_count;        // two consecutive words: queue length, consumer cursor
    uint32_t *long_cursor;
};
// mode: 0 = BIN_EMIT_LIVE (production emit), 1 = BIN_COUNT, 2 = BIN_EMIT (parity dumps).
int launch_bin(int mode, const BatchDev &b, const BinArgs &args, cudaStream_t stream);

int launch_sum_fill_counts(const uint32_t *tile_word, uint32_t n_tiles, unsigned long long *total, cudaStream_t stream);

// Results of the clip batch (PrepareClipTilesD3D11), read while a draw batch with clipped paths is
// resolved (reference: tiler.rs:114-156, propagate.cs.glsl:142-189).
struct ClipDev {
    const PathInfo *paths;         // the clip batch's path records (tile rects, dense tile offsets)
    uint32_t n_paths;
    const uint32_t *tile_word;     // fill count | backdrop of every clip tile, after its own propagate
    const uint32_t *tile_fill_end; // end of each clip tile's run in `fills`
    const PackedFill *fills;
};
// Per draw tile, written by propagate when the batch has clipped paths: 0 = not clipped by an alpha
// tile; else (dense clip tile index + 1), with TILE_CLIP_REPLACE when the (solid) draw tile takes over
// the clip tile's mask and backdrop instead of min-combining its own mask with it.
constexpr uint32_t TILE_CLIP_REPLACE = 0x80000000u;
// TileEntry.paint_ctrl flag bits (the low 24 bits are colour | ctrl).
constexpr uint32_t ENTRY_HAS_CLIP = 1u << 24, ENTRY_CLIP_REPLACE = 1u << 25, ENTRY_TEXTURED = 1u << 26;

// A paint that needs per-pixel evaluation: it samples a colour texture (TextureMetadataEntry with a colour combine
// mode, gpu_data.rs:336-344) and / or blends with something other than SrcOver. The ten RGBA16F texels of
// gpu/renderer.rs:712-763 as one device record: p0..p4 are filterParams0..4 exactly as compute_filter_params packs
// them (gpu/renderer.rs:967-1049) — text: kernel, bg, fg (w = gamma correction); radial gradient: line from + vector,
// radii + uv origin; blur: direction + support, Gaussian coefficients; colour matrix: its five columns.
constexpr uint32_t PAINT_HAS_TEXTURE = 1u;
constexpr uint32_t PAINT_COMBINE_DEST_IN = 2u; // ColorCombineMode::DestIn (else SrcIn)
struct __align__(16) PaintTexture {
    float m00, m01, m10, m11, tx, ty; // framebuffer position (pixel centre) -> normalised texture coordinate
    uint32_t filter_kind;             // PF_FILTER_*
    uint32_t flags;                   // PAINT_HAS_TEXTURE, PAINT_COMBINE_DEST_IN
    float4 p0, p1, p2, p3, p4;
    float4 base;                      // base colour, rounded through f16, not premultiplied
    uint32_t blend_mode;              // PF_BLEND_MODE_*
    uint32_t pad[3];
};

// The colour texture of a draw batch (DrawTileBatchD3D11.color_texture): one RGBA8 page in device memory.
struct ColorTexture {
    const uint8_t *pixels; // NULL: the batch has no colour texture
    size_t pitch;
    int32_t width, height;
    int32_t bottom_up;     // the page is a render target: v = 1 addresses its top row (see pf_cuda.h)
    uint32_t sampling_flags; // PF_TEXTURE_SAMPLING_FLAGS_* (TileBatchTexture.sampling_flags)
};

// clip / tile_clip: NULL unless the batch has clipped paths.
int launch_propagate(const BatchDev &b, uint32_t *tile_word, const int32_t *col_backdrop, int32_t *z_buffer,
                     const ClipDev *clip, uint32_t *tile_clip, uint32_t *tile_orig_count, cudaStream_t stream);
// (tile_orig_count, optional: the fill count of every tile before the clip was applied, for the parity dumps)

// tile_fb[t] = framebuffer tile index if the tile is non-empty, inside the framebuffer and not
// z-culled, else 0xffffffff; fb_count[fb] += 1 for every survivor.
// It also reserves each surviving tile's run in the tile-grouped fill array: tile_fill_pos[t] = run
// start, *fill_cursor += total (keep_all_fills: culled tiles keep their runs, for the parity dumps).
int launch_list_count(const BatchDev &b, const uint32_t *tile_word, const int32_t *z_buffer, uint32_t *tile_fb,
                      uint32_t *fb_count, uint32_t *tile_fill_pos, uint32_t *fill_cursor, uint32_t *path_live,
                      bool keep_all_fills, const uint32_t *run_counts, uint32_t *live_tiles, uint32_t live_capacity,
                      uint32_t *live_count, uint32_t *fb_alpha, const uint32_t *tile_clip, cudaStream_t stream);
// (live_tiles, optional: the surviving tiles as a compact list, *live_count of them, for launch_list_emit)
// (fb_alpha[fb] = 1 when a surviving tile there has fills or — tile_clip, optional — a clip mask)
// Device-side totals ([0] lines, [2] entries, [5] visible fills) and the capacities they must fit.
struct OverflowGuard {
    const uint32_t *totals;
    uint32_t line_bound, entry_bound, fill_bound;
    uint32_t *work_counter; // the fused kernel's tile counter; parked past the end on overflow
};
// Appends one TileEntry per surviving tile to its framebuffer tile's run [fb_start, fb_start + count).
int launch_list_emit(const BatchDev &b, const uint32_t *tile_fb, const uint32_t *tile_word,
                     const uint32_t *tile_fill_pos, const uint32_t *fb_start, uint32_t *fb_cursor,
                     const float4 *paints, TileEntry *entries, uint32_t capacity, const OverflowGuard &guard,
                     const ClipDev *clip, const uint32_t *tile_clip, uint2 *entry_clip, const uint32_t *live_tiles,
                     uint32_t live_capacity, const uint32_t *live_count, cudaStream_t stream);

struct CompositeArgs {
    const TileEntry *entries;   // runs in arbitrary order; the kernel sorts each by tile_index
    const uint2 *entry_clip;    // per entry {clip fill end, clip tile word}, for entries with ENTRY_HAS_CLIP (else NULL)
    const PaintTexture *paint_textures; // per paint id, for entries with ENTRY_TEXTURED (else NULL)
    ColorTexture color_texture;
    const uint8_t *gamma_lut;   // 256 x 8 L8 (textures/gamma-lut.png), NULL when the renderer was created without it
    const PackedFill *clip_fills;
    const uint32_t *fb_start, *fb_count;
    const uint32_t *fb_alpha;  // per framebuffer tile: some entry has fills or a clip mask (per-pixel work)
    uint32_t *queue;           // framebuffer tiles (row << 16 | column within the strip) that need per-pixel work,
    uint32_t *queue_count;     //   appended by k_tile_solid, consumed by k_tile_alpha; zeroed by the caller
    const PackedFill *fills;
    uint2 *queue_hdr;          // {list length, list start} of every queued tile, beside `queue` (saves a dependent load)
    cudaTextureObject_t area_lut;
    FbRect fb;
    int32_t tile_y0, tile_y1;  // tile rows composited by this renderer (strip)
    uint8_t *dest;             // local image (also dests[0])
    uint8_t *dests[8];         // local image first, then the peers' images (same layout) for the fused gather
    int n_dest;
    uint32_t dest_align_mask;  // OR of all destination addresses and the pitch (low bits: alignment)
    size_t dest_pitch;
    int32_t dest_w, dest_h;
    float4 clear_color;
    int load_dest;             // LOAD_ACTION_LOAD for batches after the first
    uint32_t *work_counter;    // device word used by the persistent warps to pull queued tiles; zeroed by the caller
    // Compact export of this strip for the other ranks (PFCudaRendererGatherFrame, tile mode; all NULL otherwise),
    // written next to the frame by the fill + tile kernels: which tiles of each 32-tile row segment have a single
    // colour and that colour, the number of queued tiles, and — slot k for queue entry k — the finished pixels of
    // every queued tile as one contiguous 1 KB block (16 rows x 64 B). `queue` then points into the export too.
    uint32_t *export_solid_color, *export_solid_mask, *export_alpha_count;
    uint8_t *export_blocks;
};
int launch_composite(const CompositeArgs &args, cudaStream_t stream);

// One peer's compact export as received by this rank (pointers into this rank's own receive slot for that peer) and
// where the peer's strip lies in the frame.
struct PullPeer {
    const uint32_t *queue;        // tile (row << 16 | column within the peer's strip) of every queued tile
    const uint32_t *alpha_count;
    const uint32_t *solid_color, *solid_mask;
    const uint8_t *blocks;
    int32_t tile_y0, tile_y1;     // the peer's strip
};
struct PullArgs {
    PullPeer peers[7];
    int n_peers;
    FbRect fb;
    uint8_t *dest;
    size_t dest_pitch;
    int32_t dest_w, dest_h;
};
// Completes the local frame with the other ranks' strips from the compact exports they pushed into this rank's
// receive slots: single-colour tiles are expanded from 4 bytes, the others copied as 1 KB blocks (all local memory).
int launch_pull_tiles(const PullArgs &args, cudaStream_t stream);

// Pushes this rank's compact export (the used part of one slot: count, queue, colours, masks, blocks) into the same
// slot of every peer's receive region over NVLink. Slot layout: the byte offsets below; `segments` = 32-tile row
// segments of this rank's strip.
struct PushArgs {
    const uint8_t *local;
    uint8_t *remote[7];
    int n_remote;
    size_t queue_off, color_off, mask_off, blocks_off;
    uint32_t segments;
};
int launch_push_export(const PushArgs &args, cudaStream_t stream);

// ---- parity-dump helpers ----
int launch_alpha_flags(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_first_fill,
                       uint8_t *fill_is_first, cudaStream_t stream);
int launch_alpha_assign(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_first_fill,
                        const uint32_t *fill_first_scan, uint32_t *tile_alpha_id, uint32_t alpha_base,
                        cudaStream_t stream);
int launch_dump_fills(uint32_t n_fills, const EmitFill *fills_emit, const uint32_t *tile_alpha_id, void *out,
                      cudaStream_t stream);
int launch_dump_tile_flags(uint32_t n_tiles, const uint32_t *tile_word, uint32_t *flags, cudaStream_t stream);
// tile_clip / clip_tile_word / clip_alpha_id: NULL unless the batch has clipped paths.
int launch_dump_tiles(const BatchDev &b, const uint32_t *tile_word, const uint32_t *tile_alpha_id,
                      const uint32_t *flags, const uint32_t *pos, void *out, const uint32_t *tile_clip,
                      const uint32_t *clip_tile_word, const uint32_t *clip_alpha_id, cudaStream_t stream);
int launch_dump_clip_flags(uint32_t n_tiles, const uint32_t *tile_clip, uint32_t *flags, cudaStream_t stream);
int launch_dump_clips(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_alpha_id, const uint32_t *tile_clip,
                      const uint32_t *clip_tile_word, const uint32_t *clip_alpha_id, const uint32_t *flags,
                      const uint32_t *pos, void *out, cudaStream_t stream);
int launch_alpha_masks(uint32_t n_tiles, const uint32_t *tile_word, const uint32_t *tile_fill_pos,
                       const uint32_t *tile_alpha_id, const PackedFill *fills, cudaTextureObject_t area_lut,
                       float *out, cudaStream_t stream);

} // namespace pf
