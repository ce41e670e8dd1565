// pathfinder_b200/csrc/renderer.cu — host driver of the CUDA pipeline and the renderer half of the
// C ABI (include/pf_cuda.h).
//
// Replaces RendererD3D11::{upload_scene, prepare_tiles, draw_tiles} (renderer/src/gpu/d3d11/
// renderer.rs:236-241,427-542,711-782) and Renderer::{begin_scene, render_command, end_scene}
// (renderer/src/gpu/renderer.rs:350-460). Differences by design:
//   * the first frame of a batch is count -> scan -> emit with exact allocation; later frames keep the
//     counts on the device, size grids from the previous frame and verify once at the end, so there
//     is no overflow/retry loop per stage (the reference re-runs dice/bin with doubled buffers,
//     d3d11/renderer.rs:463-499) and at most one host wait per frame (none with deferred verification);
//   * occlusion culling happens before fill emission; fill and tile are one kernel (the alpha mask
//     never reaches HBM);
//   * per-framebuffer-tile lists are runs from a count + scan, rank-sorted inside the fused kernel,
//     instead of per-tile linked lists with an insertion sort;
//   * clip paths (one level) are a second batch through the same stages, resolved in propagate.
#include <cuda.h>
#include <cuda_fp16.h>
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <vector>

#include "../../include/pf_cuda.h"
#include "common.cuh"
#include "kernels.cuh"
#include "scan.cuh"

namespace pf {

thread_local std::string g_last_error;

void set_last_error(const std::string &msg) { g_last_error = msg; }

template <typename T>
struct PinnedBuffer {
    T *ptr = nullptr;
    size_t capacity = 0;
    ~PinnedBuffer() {
        if (ptr) cudaFreeHost(ptr);
    }
    void ensure(size_t n) {
        if (n <= capacity) return;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        size_t want = n + n / 4 + 64;
        PF_CUDA_CHECK(cudaMallocHost((void **)&ptr, want * sizeof(T)));
        capacity = want;
    }
};

// A typed window into PFCudaRenderer::zeroed.
template <typename T>
struct ZeroedView {
    T *ptr = nullptr;
};

struct SceneSegments {
    DeviceBuffer<float2> points;
    DeviceBuffer<uint2> indices;
    size_t n_points = 0, n_indices = 0;
};

// What the renderer keeps about the batch it last uploaded, so that a batch with the same
// content_key is rendered without any host-side rebuild or upload.
struct BatchCache {
    bool valid = false;
    uint64_t key = 0;
    uint64_t scene_generation = 0, paint_generation = 0;
    int32_t strip_y0 = 0, strip_y1 = 0;
    BatchDev batch{};
    bool has_initial_backdrops = false;
    bool has_clips = false; // the batch has clipped paths (resolved against r->clip)
    ColorTexture color_texture{nullptr, 0, 0, 0, 0, 0}; // DrawTileBatchD3D11.color_texture, resolved to its page
    bool counts_valid = false; // n_lines / n_fills / n_entries hold the last frame's totals
    uint32_t n_lines = 0, n_fills = 0, n_entries = 0, n_visible_fills = 0;
    uint32_t command_paths = 0, command_segments = 0; // as sent (before strip culling)
};

constexpr uint32_t LONG_QUEUE_CAPACITY = 1u << 18; // more long lines than this fall back to the per-thread walk

// Bounds of the batch whose totals have not been checked yet (deferred verification).
struct PendingVerify {
    bool active = false;
    uint32_t line_bound = 0, fill_bound = 0, entry_bound = 0, emit_bound = 0;
    int launches = 0;
    int batches_drawn_before = 0;
};

struct StageTimer {
    cudaEvent_t ev[8];
    bool created = false;
    void create() {
        if (created) return;
        for (auto &e : ev) PF_CUDA_CHECK(cudaEventCreate(&e));
        created = true;
    }
    void destroy() {
        if (!created) return;
        for (auto &e : ev) cudaEventDestroy(e);
        created = false;
    }
};

} // namespace pf

using namespace pf;

struct PFCudaDevice {
    int ordinal;
};

struct PFCudaRenderer {
    int ordinal = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    PFCudaRendererOptions options{};
    size_t bytes_allocated = 0;

    // Destination image.
    DeviceBuffer<uint8_t> dest_owned;
    uint8_t *dest = nullptr;
    size_t dest_pitch = 0;
    bool dest_external = false;
    int32_t dest_external_rows = 0; // rows the external / peer destinations were installed for

    // Area LUT texture (textures/area-lut.png; renderer/src/gpu/renderer.rs:207-214).
    cudaArray_t lut_array = nullptr;
    cudaTextureObject_t lut_tex = 0;

    // Scene-resident data (UploadSceneD3D11).
    SceneSegments draw_segments, clip_segments;
    bool has_scene = false;

    // Paint table (UploadTextureMetadata), base colours rounded through f16.
    DeviceBuffer<float4> paints;
    size_t n_paints = 0;
    // Paints that sample a colour texture (colour combine mode SrcIn): per-paint record for the compositing kernel,
    // and which paints those are (PathInfo gets PATH_TEXTURED).
    DeviceBuffer<PaintTexture> paint_textures;
    std::vector<uint8_t> paint_is_textured;
    bool any_textured_paint = false;
    DeviceBuffer<uint8_t> gamma_lut; // textures/gamma-lut.png, 256 x 8 L8 (only the text filter reads it)
    bool has_gamma_lut = false;

    // Texture pages (AllocateTexturePage) and render targets (DeclareRenderTarget / PushRenderTarget /
    // PopRenderTarget; renderer/src/gpu/renderer.rs:462-520,1151-1209). Draw batches go to the render target on
    // top of the stack, or to the destination image when the stack is empty.
    struct TexturePage {
        DeviceBuffer<uint8_t> pixels; // RGBA8, pitch = 4 * width
        int32_t width = 0, height = 0;
        bool is_render_target = false;
    };
    struct RenderTargetSlot {
        bool declared = false;
        uint32_t page = 0;
        PFRectI rect{{0, 0}, {0, 0}};
        int batches_drawn = 0; // this frame: the first batch starts from transparent black
    };
    std::vector<std::unique_ptr<TexturePage>> pages;
    std::vector<RenderTargetSlot> render_targets;
    std::vector<uint32_t> target_stack;

    // Strip partition.
    int32_t strip_y0 = 0, strip_y1 = 0;
    // Peers' copies of the frame (IPC-mapped), written by the fused gather in k_composite.
    uint8_t *peer_dest[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void *peer_base[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n_peers = 0;

    // Frame assembly across GPUs (PFCudaRendererGather*): NCCL communicator, its stream, and the two events that
    // order it after this frame's compositing and the next frame's compositing after it.
    void *gather_comm = nullptr; // ncclComm_t
    int gather_rank = 0, gather_world = 0;
    cudaStream_t gather_stream = nullptr;
    cudaEvent_t gather_ready = nullptr, gather_done = nullptr;
    bool gather_in_flight = false, gather_last_was_tiles = false;
    // Tile mode (PF_CUDA_GATHER_MODE_TILES): every rank owns a receive region that its peers map through CUDA IPC,
    // with one slot per (source rank, frame parity); frame i of rank g lands in slot (g, i % 2) of every other rank —
    // and first, written by rank g's own fill + tile kernels, in slot (g, i % 2) of its own region.
    // Layout of one slot (the same everywhere, sized for the tallest strip): [alpha count, 256 B][queue: u32 per tile]
    // [solid colour: u32 per tile slot][solid mask: u32 per segment][blocks: 1 KB per tile].
    int gather_mode = 0;
    DeviceBuffer<uint8_t> export_region;
    size_t export_buffer_bytes = 0, export_queue_off = 0, export_color_off = 0, export_mask_off = 0, export_blocks_off = 0;
    uint8_t *peer_export[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // by rank (own: local pointer)
    void *peer_export_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t gather_frame_serial = 0;  // frames gathered so far: selects the export buffer
    bool frame_exported = false;       // this frame's (single) destination batch wrote the current export buffer
    int main_batches_this_frame = 0;
    uint32_t *barrier_word = nullptr;  // device word all-reduced as the barrier
    cudaEvent_t gather_barrier_done = nullptr; // tile mode: every rank's push of the gathered frame has landed
    cudaEvent_t push_done[2] = {nullptr, nullptr}; // tile mode: the push that read the local export slot of that parity
    bool push_recorded[2] = {false, false};

    // Per-batch device buffers.
    DeviceBuffer<uint8_t> batch_meta; // PathInfo[P] + 3 search arrays
    PinnedBuffer<uint8_t> batch_meta_host;
    cudaEvent_t meta_copied = nullptr;
    DeviceBuffer<uint32_t> seg_line_offset;  // [S]
    DeviceBuffer<float4> lines;
    DeviceBuffer<uint32_t> line_path;
    DeviceBuffer<uint32_t> line_fill_offset; // [L]
    // Everything a frame needs zeroed lives in one allocation and is cleared by one memset (carve_zeroed).
    DeviceBuffer<uint32_t> zeroed;
    ZeroedView<uint32_t> counters;  // device-side totals: [0]=lines [1]=fills [2]=entries [3]=alpha tiles [4]=dump tiles
                                    // [5]=visible fills [6,7]/[13,14]=long-line queues [8,9]=u64 scratch [10]=queued framebuffer tiles
                                    // [12]=tile counter [15]=surviving tiles
    ZeroedView<uint32_t> path_live; // per path: some tile with fills survived the z-cull
    ZeroedView<uint32_t> tile_word;
    ZeroedView<int32_t> col_backdrop;
    ZeroedView<int32_t> z_buffer;
    ZeroedView<uint32_t> fb_count, fb_cursor, fb_alpha;
    DeviceBuffer<uint32_t> tile_fill_pos, tile_first_fill, tile_fb, tile_pos, tile_alpha_id;
    DeviceBuffer<int32_t> col_backdrop_init;
    DeviceBuffer<uint32_t> long_queue; // lines walked by whole warps (k_bin_long)
    // Clip batch results (PrepareClipTilesD3D11), kept until the draw batch of the same frame has used them.
    struct ClipStage {
        bool valid = false;
        uint32_t n_paths = 0;
        DeviceBuffer<uint8_t> meta;              // PathInfo[n_paths]
        DeviceBuffer<uint32_t> tile_word, tile_fill_end;
        DeviceBuffer<PackedFill> fills;
        // parity dumps: alpha tile ids of the clip tiles (they come first in SequentialExecutor order)
        // and the clip paths' fills as records
        DeviceBuffer<uint32_t> tile_alpha_id;
        DeviceBuffer<uint8_t> fill_records;
        uint32_t n_fills = 0, n_alpha = 0;
    } clip;
    DeviceBuffer<uint32_t> tile_orig;  // parity dumps with clips: fill count of every draw tile before the clip
    DeviceBuffer<uint32_t> live_tiles; // steady state: the tiles that survived the z-cull, compacted
    DeviceBuffer<uint32_t> tile_clip;  // per draw tile: clip tile reference (batches with clipped paths)
    DeviceBuffer<uint2> entry_clip;    // per list entry: {clip fill end, clip tile word}
    DeviceBuffer<PackedFill> fills;
    DeviceBuffer<EmitFill> fills_emit;
    DeviceBuffer<uint32_t> fb_start;
    DeviceBuffer<uint32_t> tile_queue; // framebuffer tiles that need per-pixel compositing (k_tile_solid -> k_tile_alpha)
    DeviceBuffer<uint2> tile_queue_hdr; // their list headers, same index
    DeviceBuffer<TileEntry> entries;
    PinnedBuffer<uint32_t> counters_host;
    ScanScratch scan_scratch;
    // debug / dump scratch
    DeviceBuffer<uint8_t> fill_is_first;
    DeviceBuffer<uint32_t> fill_first_scan;
    DeviceBuffer<uint8_t> dump_out;

    BatchCache cache;
    PendingVerify pending;
    cudaEvent_t verify_event = nullptr;
    bool deferred_verify = false;
    // Host-to-device copies from caller-owned arrays (payload_persists) that no host wait has covered yet.
    bool borrowed_copies_pending = false;
    bool uploaded_scene_this_frame = false;
    uint64_t scene_generation = 0, paint_generation = 0;
    uint64_t paint_key = 0;
    bool always_size = false; // debugging aid: read every count back (three syncs per batch)

    // State of the last batch (for dumps and stats).
    BatchDev last_batch{};
    uint32_t last_lines = 0, last_fills = 0, last_entries = 0, last_alpha_tiles = 0;
    bool last_alpha_ids_valid = false;
    FbRect last_fb{0, 0, 0, 0};

    bool in_scene = false;
    bool debug_lists = false;
    bool timing = false;
    int batches_drawn = 0;
    PFCudaRenderStats stats{};
    PFCudaRenderTime times{};       // the current frame
    PFCudaRenderTime times_total{}; // every batch since timing was switched on
    uint32_t times_batches = 0;
    StageTimer timer;

    // scene.view_box(): what process_line_segment clips to (renderer/src/tiler.rs:194). Defaults to
    // the destination rect, which is what the demo sets (demo/common/src/lib.rs:910-914).
    bool has_view_box = false;
    ViewBox view_box{0, 0, 0, 0};
    // SceneSink.last_scene for PFSceneBuildAndRenderCuda (renderer/src/scene.rs:384-396).
    PFSceneSinkState sink_state{0, 0, 0};

    void bind_device() { PF_CUDA_CHECK(cudaSetDevice(ordinal)); }
};

namespace {

template <typename T>
void track(PFCudaRenderer *r, DeviceBuffer<T> &b) {
    b.bytes_allocated = &r->bytes_allocated;
}

void setup_tracking(PFCudaRenderer *r) {
    track(r, r->dest_owned);
    track(r, r->draw_segments.points);
    track(r, r->draw_segments.indices);
    track(r, r->clip_segments.points);
    track(r, r->clip_segments.indices);
    track(r, r->paints);
    track(r, r->paint_textures);
    track(r, r->gamma_lut);
    track(r, r->batch_meta);
    track(r, r->seg_line_offset);
    track(r, r->lines);
    track(r, r->line_path);
    track(r, r->line_fill_offset);
    track(r, r->zeroed);
    track(r, r->tile_fill_pos);
    track(r, r->tile_first_fill);
    track(r, r->tile_fb);
    track(r, r->tile_pos);
    track(r, r->tile_alpha_id);
    track(r, r->col_backdrop_init);
    track(r, r->long_queue);
    track(r, r->clip.meta);
    track(r, r->clip.tile_word);
    track(r, r->clip.tile_fill_end);
    track(r, r->clip.fills);
    track(r, r->clip.tile_alpha_id);
    track(r, r->clip.fill_records);
    track(r, r->tile_orig);
    track(r, r->live_tiles);
    track(r, r->tile_clip);
    track(r, r->entry_clip);
    track(r, r->fills);
    track(r, r->fills_emit);
    track(r, r->fb_start);
    track(r, r->tile_queue);
    track(r, r->tile_queue_hdr);
    track(r, r->entries);
    track(r, r->scan_scratch.control);
    track(r, r->scan_scratch.status);
    track(r, r->fill_is_first);
    track(r, r->fill_first_scan);
    track(r, r->dump_out);
}

void allocate_dest(PFCudaRenderer *r) {
    if (r->dest_external) return;
    int w = r->options.dest_size.x, h = r->options.dest_size.y;
    if (w <= 0 || h <= 0) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "dest_size must be positive");
    r->dest_owned.ensure((size_t)w * h * 4);
    r->dest = r->dest_owned.ptr;
    r->dest_pitch = (size_t)w * 4;
}

// Where draw batches currently go: the render target on top of the stack, or the destination image
// (Renderer::draw_render_target / main_viewport, renderer/src/gpu/renderer.rs:1151-1209).
struct DrawTarget {
    uint8_t *pixels;
    size_t pitch;
    int32_t width, height;
    bool is_main;
    int *batches_drawn;
};
DrawTarget current_target(PFCudaRenderer *r) {
    if (r->target_stack.empty())
        return DrawTarget{r->dest, r->dest_pitch, r->options.dest_size.x, r->options.dest_size.y, true, &r->batches_drawn};
    PFCudaRenderer::RenderTargetSlot &slot = r->render_targets[r->target_stack.back()];
    PFCudaRenderer::TexturePage &page = *r->pages[slot.page];
    const size_t pitch = (size_t)page.width * 4;
    return DrawTarget{page.pixels.ptr + (size_t)slot.rect.origin.y * pitch + (size_t)slot.rect.origin.x * 4, pitch,
                      slot.rect.lower_right.x - slot.rect.origin.x, slot.rect.lower_right.y - slot.rect.origin.y, false,
                      &slot.batches_drawn};
}

// round_out(view_box / 16): the framebuffer tile rect (renderer/src/builder.rs:949-953) of the current draw target.
FbRect target_tile_rect(const DrawTarget &t) {
    FbRect fb;
    fb.min_x = 0;
    fb.min_y = 0;
    fb.max_x = (t.width + PF_TILE_WIDTH - 1) / PF_TILE_WIDTH;
    fb.max_y = (t.height + PF_TILE_HEIGHT - 1) / PF_TILE_HEIGHT;
    return fb;
}
// ... of the destination image.
FbRect framebuffer_tile_rect(const PFCudaRenderer *r) {
    FbRect fb;
    fb.min_x = 0;
    fb.min_y = 0;
    fb.max_x = (r->options.dest_size.x + PF_TILE_WIDTH - 1) / PF_TILE_WIDTH;
    fb.max_y = (r->options.dest_size.y + PF_TILE_HEIGHT - 1) / PF_TILE_HEIGHT;
    return fb;
}

// Local image first, then — for the destination image only — the peers' images (fused all-gather, see composite.cu).
void fill_destinations(const PFCudaRenderer *r, const DrawTarget &target, CompositeArgs &ca) {
    ca.dest = target.pixels;
    ca.dests[0] = target.pixels;
    ca.n_dest = 1;
    ca.dest_pitch = target.pitch;
    ca.dest_w = target.width;
    ca.dest_h = target.height;
    uintptr_t bits = (uintptr_t)target.pixels | (uintptr_t)target.pitch;
    for (int i = 0; target.is_main && i < r->n_peers && ca.n_dest < 8; i++) {
        ca.dests[ca.n_dest++] = r->peer_dest[i];
        bits |= (uintptr_t)r->peer_dest[i];
    }
    ca.dest_align_mask = (uint32_t)(bits & 0xffffffffu);
}

// The previous frame may still be on its way to the other ranks from the destination buffer (PFCudaRendererGatherFrame):
// whatever writes the destination next is ordered after it.
void wait_for_gather(PFCudaRenderer *r, cudaStream_t st) {
    if (!r->gather_in_flight) return;
    PF_CUDA_CHECK(cudaStreamWaitEvent(st, r->gather_done, 0));
    r->gather_in_flight = false;
}
// What the next compositing of the destination has to wait for. FRAME mode: the whole gather (it sends from and
// receives into the buffer the compositing writes). TILES mode: nothing of the previous gather — its push reads the
// other export slot, its pull kernel writes the OTHER ranks' strips — only the push two frames back, which read the
// local export slot this frame is about to overwrite (run_pipeline waits for that one).
void wait_for_gather_before_compositing(PFCudaRenderer *r, cudaStream_t st, bool last_was_tiles) {
    if (!r->gather_in_flight || last_was_tiles) return; // (gather_in_flight stays set in tile mode: readers still wait for the pull)
    wait_for_gather(r, st);
}

float4 clear_color(const PFCudaRenderer *r) {
    // Renderer::clear_color_for_draw_operation (gpu/renderer.rs): background colour if set,
    // otherwise transparent black.
    if (r->options.flags & PF_RENDERER_OPTIONS_FLAGS_HAS_BACKGROUND_COLOR) {
        // (clamped: the compositing kernels pack without saturating — every colour they blend lies in [0, 1])
        const PFColorF &c = r->options.background_color;
        auto unit = [](float v) { return v >= 0.0f ? (v <= 1.0f ? v : 1.0f) : 0.0f; }; // NaN -> 0
        return make_float4(unit(c.r), unit(c.g), unit(c.b), unit(c.a));
    }
    return make_float4(0, 0, 0, 0);
}

uint32_t read_counter(PFCudaRenderer *r, int index) {
    r->counters_host.ensure(16);
    PF_CUDA_CHECK(cudaMemcpyAsync(r->counters_host.ptr, r->counters.ptr, 8 * sizeof(uint32_t),
                                  cudaMemcpyDeviceToHost, r->stream));
    PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    r->stats.host_sync_count++;
    return r->counters_host.ptr[index];
}

struct HostTimer {
    const char *what;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit HostTimer(const char *w) : what(w) {}
    ~HostTimer() {
        static const bool on = getenv("PF_HOST_TIMING") != nullptr;
        if (on)
            fprintf(stderr, "%s: %.3f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

void upload_segments(PFCudaRenderer *r, SceneSegments &dst, const PFSegmentsD3D11 &src, bool payload_persists) {
    HostTimer timer("upload_segments");
    LapTimer laps;
    // Every index must address its 2, 3 or 4 points inside `points` (SegmentsD3D11::add_path, builder.rs:807-841):
    // dice reads them unchecked.
    {
        std::atomic<bool> bad{false};
        const PFSegmentIndicesD3D11 *idx = src.indices;
        const size_t n_points = src.point_count;
        parallel_ranges(src.index_count, 65536, [&](size_t begin, size_t end) {
            bool b = false;
            for (size_t i = begin; i < end; i++) {
                const uint32_t flags = idx[i].flags;
                const size_t last = (size_t)idx[i].first_point_index + ((flags & 0x40000000u) ? 3u : (flags & 0x80000000u) ? 2u : 1u);
                b |= last >= n_points;
            }
            if (b) bad.store(true, std::memory_order_relaxed);
        });
        if (bad.load()) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "UploadSceneD3D11: a segment index points past the point array");
    }
    dst.n_points = src.point_count;
    dst.n_indices = src.index_count;
    dst.points.ensure(src.point_count + 4);
    dst.indices.ensure(src.index_count + 1);
    laps.lap("segments: ensure");
    if (src.point_count)
        PF_CUDA_CHECK(cudaMemcpyAsync(dst.points.ptr, src.points, src.point_count * sizeof(float2),
                                      cudaMemcpyHostToDevice, r->stream));
    laps.lap("segments: enqueue points");
    if (src.index_count)
        PF_CUDA_CHECK(cudaMemcpyAsync(dst.indices.ptr, src.indices, src.index_count * sizeof(uint2),
                                      cudaMemcpyHostToDevice, r->stream));
    laps.lap("segments: enqueue indices");
    if (src.point_count == 0 && src.index_count == 0) return;
    if (payload_persists) {
        r->borrowed_copies_pending = true; // covered by the frame's verification wait (or EndScene)
    } else {
        // The payload is borrowed for this call only and may be pageable: wait for the copies.
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    }
    r->stats.h2d_bytes += src.point_count * sizeof(float2) + src.index_count * sizeof(uint2);
}

void upload_texture_metadata(PFCudaRenderer *r, const PFTextureMetadataEntry *entries, size_t n) {
    std::vector<float4> table(n);
    std::vector<PaintTexture> textures(n);
    r->paint_is_textured.assign(n, 0);
    r->any_textured_paint = false;
    auto half_round = [](float v) { return __half2float(__float2half_rn(v)); }; // the metadata texture is RGBA16F
    for (size_t i = 0; i < n; i++) {
        const PFTextureMetadataEntry &e = entries[i];
        if (e.color_0_combine_mode > PF_COLOR_COMBINE_MODE_DEST_IN)
            throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "unknown colour combine mode");
        const bool textured = e.color_0_combine_mode != PF_COLOR_COMBINE_MODE_NONE;
        if (e.blend_mode > PF_BLEND_MODE_LUMINOSITY) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "unknown blend mode");
        if (e.filter.kind > PF_FILTER_COLOR_MATRIX) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "unknown filter kind");
        if (e.filter.kind != PF_FILTER_NONE && !textured)
            throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "a filter on a paint without a colour texture");
        // ColorU::to_f32 (color/src/lib.rs:70-73) then f16 (gpu/renderer.rs:726-729).
        const float s = 1.0f / 255.0f;
        float c[4] = {(float)e.base_color.r * s, (float)e.base_color.g * s, (float)e.base_color.b * s,
                      (float)e.base_color.a * s};
        for (float &v : c) v = half_round(v);
        table[i] = make_float4(c[0], c[1], c[2], c[3]);
        PaintTexture &pt = textures[i];
        memset(&pt, 0, sizeof(pt));
        pt.base = table[i];
        pt.blend_mode = e.blend_mode;
        if (textured) {
            // gpu/renderer.rs:712-763: transform, filter parameters and colours all pass through f16.
            const PFTransform2F &t = e.color_0_transform;
            pt.m00 = half_round(t.matrix.m00), pt.m01 = half_round(t.matrix.m01);
            pt.m10 = half_round(t.matrix.m10), pt.m11 = half_round(t.matrix.m11);
            pt.tx = half_round(t.vector.x), pt.ty = half_round(t.vector.y);
            pt.filter_kind = e.filter.kind;
            pt.flags = PAINT_HAS_TEXTURE | (e.color_0_combine_mode == PF_COLOR_COMBINE_MODE_DEST_IN ? PAINT_COMBINE_DEST_IN : 0u);
            const float *fp = e.filter.params;
            auto h4 = [&](float x, float y, float z, float w) { return make_float4(half_round(x), half_round(y), half_round(z), half_round(w)); };
            // compute_filter_params (gpu/renderer.rs:967-1049)
            if (e.filter.kind == PF_FILTER_TEXT) {
                // p0 = kernel (or zero), p1 = bg, p2 = fg + gamma flag
                const bool gamma = (e.filter.flags & PF_FILTER_FLAG_TEXT_GAMMA_CORRECTION) != 0;
                if (e.filter.flags & PF_FILTER_FLAG_TEXT_HAS_KERNEL) pt.p0 = h4(fp[8], fp[9], fp[10], fp[11]);
                pt.p1 = h4(fp[4], fp[5], fp[6], 0.0f);
                pt.p2 = h4(fp[0], fp[1], fp[2], gamma ? 1.0f : 0.0f);
                if (gamma && !r->has_gamma_lut)
                    throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "a text filter asks for gamma correction but the renderer was created without the gamma LUT");
            } else if (e.filter.kind == PF_FILTER_RADIAL_GRADIENT) {
                // params = line.from, line.to, radii, uv_origin -> p0 = from, vector; p1 = radii, uv origin
                pt.p0 = h4(fp[0], fp[1], fp[2] - fp[0], fp[3] - fp[1]);
                pt.p1 = h4(fp[4], fp[5], fp[6], fp[7]);
            } else if (e.filter.kind == PF_FILTER_BLUR) {
                const float sigma = fp[0];
                if (!(sigma > 0.0f) || sigma > 1024.0f) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "blur sigma out of range");
                const float sigma_inv = 1.0f / sigma;
                const float gx = 0.3989422804014327f * sigma_inv; // SQRT_2_PI_INV
                const float gy = expf(-0.5f * sigma_inv * sigma_inv);
                const bool vertical = (e.filter.flags & PF_FILTER_FLAG_BLUR_Y) != 0;
                pt.p0 = h4(vertical ? 0.0f : 1.0f, vertical ? 1.0f : 0.0f, ceilf(1.5f * sigma) * 2.0f, 0.0f);
                pt.p1 = h4(gx, gy, gy * gy, 0.0f);
            } else if (e.filter.kind == PF_FILTER_COLOR_MATRIX) {
                pt.p0 = h4(fp[0], fp[1], fp[2], fp[3]);
                pt.p1 = h4(fp[4], fp[5], fp[6], fp[7]);
                pt.p2 = h4(fp[8], fp[9], fp[10], fp[11]);
                pt.p3 = h4(fp[12], fp[13], fp[14], fp[15]);
                pt.p4 = h4(fp[16], fp[17], fp[18], fp[19]);
            }
        }
        if (textured || e.blend_mode != PF_BLEND_MODE_SRC_OVER) {
            r->paint_is_textured[i] = 1; // evaluated per pixel by the GENERAL variant of k_tile_alpha
            r->any_textured_paint = true;
        }
    }
    r->n_paints = n;
    r->paints.ensure(n + 1);
    r->paint_textures.ensure(n + 1);
    if (n) {
        PF_CUDA_CHECK(cudaMemcpyAsync(r->paints.ptr, table.data(), n * sizeof(float4), cudaMemcpyHostToDevice,
                                      r->stream));
        if (r->any_textured_paint)
            PF_CUDA_CHECK(cudaMemcpyAsync(r->paint_textures.ptr, textures.data(), n * sizeof(PaintTexture),
                                          cudaMemcpyHostToDevice, r->stream));
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        r->stats.h2d_bytes += n * sizeof(float4);
    }
}

// Builds the per-path device records of a batch (host part of "bound") and uploads them. Replaces
// the offsets TileBatchDataD3D11::push assigns (renderer/src/builder.rs:663-720) + bound.cs.glsl.
void upload_batch_metadata(PFCudaRenderer *r, const PFTileBatchDataD3D11 &batch, const FbRect &fb,
                           int32_t strip_y0, int32_t strip_y1, BatchDev &b, bool &has_initial_backdrops,
                           const SceneSegments &segments, bool is_clip_batch) {
    HostTimer timer("upload_batch_metadata");
    LapTimer laps;
    cudaStream_t st = r->stream;
    const uint32_t P = batch.path_count;
    const PFPrepareTilesInfoD3D11 &info = batch.prepare_info;
    // Coarse search tables (see CoarseIndex); sized once pass 1 has counted the kept tiles / columns.
    constexpr int SEG_SHIFT = 5, TILE_SHIFT = 7, COL_SHIFT = 5;
    const uint32_t n_paints = (uint32_t)r->n_paints;
    // A path without a tile in this strip contributes nothing (every fill is culled by add_fill,
    // every backdrop adjustment by the rect test): it is dropped, segments included, so dice / bin
    // / propagate only see the paths that reach the strip.
    // Strip restriction: rows above the strip feed the column backdrops exactly like rows above the
    // path rect do in the reference (builder.rs:609-612); rows below are ignored.
    auto strip_rect = [&](uint32_t i, int32_t &w, int32_t &h) {
        const PFRectI &tr = info.propagate_metadata[i].tile_rect;
        const int32_t min_y = std::max(tr.origin.y, strip_y0), max_y = std::min(tr.lower_right.y, strip_y1);
        w = tr.lower_right.x - tr.origin.x, h = max_y - min_y;
        return w > 0 && h > 0;
    };
    auto segments_of = [&](uint32_t i) {
        const uint32_t seg_begin = info.dice_metadata[i].first_batch_segment_index;
        const uint32_t seg_end = i + 1 < P ? info.dice_metadata[i + 1].first_batch_segment_index : batch.segment_count;
        return seg_end - seg_begin;
    };
    // Pass 1 (parallel): what the kept paths of each chunk add to the running offsets (integer only);
    // pass 2 (same chunks) starts from the prefix over chunks — a two-level scan.
    struct ChunkSums {
        uint64_t kept = 0, segments = 0, tiles = 0, columns = 0;
        uint64_t pad[4]; // one cache line per chunk
    };
    const size_t chunks = chunk_count(P, 8192);
    std::vector<ChunkSums> sums(chunks + 1);
    // The payload is not trusted: a stale or malformed batch (say DrawTilesD3D11 after a smaller UploadSceneD3D11)
    // must come back as a status, not as an out-of-bounds read on the device that poisons the CUDA context.
    // 1: segment indices not monotonic / past segment_count, 2: segments outside the uploaded scene,
    // 3: inverted tile rect, 4: clip path index outside the clip batch.
    std::atomic<int> malformed{0};
    const uint32_t clip_paths = is_clip_batch ? 0u : (r->clip.valid ? r->clip.n_paths : 0u);
    parallel_chunks(P, chunks, [&](size_t chunk, size_t begin, size_t end) {
        ChunkSums sum;
        for (size_t i = begin; i < end; i++) {
            const uint32_t seg_begin = info.dice_metadata[i].first_batch_segment_index;
            const uint32_t seg_end = i + 1 < P ? info.dice_metadata[i + 1].first_batch_segment_index : batch.segment_count;
            if (seg_end < seg_begin || seg_end > batch.segment_count) malformed.store(1, std::memory_order_relaxed);
            else if ((uint64_t)info.dice_metadata[i].first_global_segment_index + (seg_end - seg_begin) > segments.n_indices)
                malformed.store(2, std::memory_order_relaxed);
            const PFRectI &tr = info.propagate_metadata[i].tile_rect;
            if (tr.lower_right.x < tr.origin.x || tr.lower_right.y < tr.origin.y) malformed.store(3, std::memory_order_relaxed);
            const uint32_t clip_index = info.propagate_metadata[i].clip_path_index;
            if (clip_index != 0xffffffffu && clip_index >= clip_paths) malformed.store(4, std::memory_order_relaxed);
            if (malformed.load(std::memory_order_relaxed)) continue;
            int32_t w, h;
            if (!strip_rect((uint32_t)i, w, h)) continue;
            sum.kept++;
            sum.segments += segments_of((uint32_t)i);
            sum.tiles += (uint64_t)w * (uint64_t)h;
            sum.columns += (uint64_t)w;
        }
        sums[chunk + 1] = sum;
    });
    for (size_t c = 1; c <= chunks; c++) { // exclusive prefix: sums[c] = totals of the chunks before c
        sums[c].kept += sums[c - 1].kept;
        sums[c].segments += sums[c - 1].segments;
        sums[c].tiles += sums[c - 1].tiles;
        sums[c].columns += sums[c - 1].columns;
    }
    switch (malformed.load()) {
    case 1: throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "batch: first_batch_segment_index is not monotonic or exceeds segment_count");
    case 2: throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "batch: a path's segments lie outside the uploaded scene (stale batch?)");
    case 3: throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "batch: inverted tile rect");
    case 4: throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "batch: clip_path_index outside the clip batch");
    default: break;
    }
    if (sums[chunks].tiles >= 0xfffffff0ull) throw Error(PF_CUDA_ERROR_UNSUPPORTED, "more than 2^32 bbox tiles in one batch");
    const uint32_t kept = (uint32_t)sums[chunks].kept, n_segments = (uint32_t)sums[chunks].segments,
                   n_tiles = (uint32_t)sums[chunks].tiles, n_cols = (uint32_t)sums[chunks].columns;
    laps.lap("meta pass 1");

    // Host staging / device layout: [P records][3 x (P + 1) offsets][segment table][tile table][column table].
    const size_t seg_table_n = ((size_t)n_segments >> SEG_SHIFT) + 2, tile_table_n = ((size_t)n_tiles >> TILE_SHIFT) + 2,
                 col_table_n = ((size_t)n_cols >> COL_SHIFT) + 2;
    const size_t base_bytes = (size_t)P * sizeof(PathInfo) + 3 * (size_t)(P + 1) * sizeof(uint32_t);
    const size_t meta_bytes = base_bytes + (seg_table_n + tile_table_n + col_table_n) * sizeof(uint32_t);
    if (r->meta_copied) PF_CUDA_CHECK(cudaEventSynchronize(r->meta_copied));
    r->batch_meta_host.ensure(meta_bytes + 64);
    uint8_t *hbase = r->batch_meta_host.ptr;
    PathInfo *h_paths = reinterpret_cast<PathInfo *>(hbase);
    uint32_t *h_seg_first = reinterpret_cast<uint32_t *>(hbase + (size_t)P * sizeof(PathInfo));
    uint32_t *h_tile_off = h_seg_first + (P + 1);
    uint32_t *h_col_off = h_tile_off + (P + 1);
    h_seg_first[kept] = n_segments;
    h_tile_off[kept] = n_tiles;
    h_col_off[kept] = n_cols;
    // Pass 2 (parallel): running offsets and the 48-byte records of the kept paths.
    std::atomic<bool> bad_paint_flag{false};
    parallel_chunks(P, chunks, [&](size_t chunk, size_t begin, size_t end) {
        bool bad = false;
        uint32_t k = (uint32_t)sums[chunk].kept, seg_first = (uint32_t)sums[chunk].segments,
                 tile_off = (uint32_t)sums[chunk].tiles, col_off = (uint32_t)sums[chunk].columns;
        for (size_t i = begin; i < end; i++) {
            const PFTilePathInfoD3D11 &tp = info.tile_path_info[i];
            bad |= !is_clip_batch && tp.color >= n_paints; // clip paths have no paint
            int32_t w, h;
            if (!strip_rect((uint32_t)i, w, h)) continue;
            const PFPropagateMetadataD3D11 &pm = info.propagate_metadata[i];
            const PFDiceMetadataD3D11 &dm = info.dice_metadata[i];
            PathInfo pi;
            pi.min_x = pm.tile_rect.origin.x;
            pi.max_x = pm.tile_rect.lower_right.x;
            pi.min_y = std::max(pm.tile_rect.origin.y, strip_y0);
            pi.max_y = std::min(pm.tile_rect.lower_right.y, strip_y1);
            pi.tile_offset = tile_off;
            pi.col_offset = col_off;
            pi.seg_batch_first = seg_first;
            pi.seg_global_first = dm.first_global_segment_index;
            pi.global_path_id = dm.global_path_id;
            pi.paint_ctrl = (uint32_t)tp.color | ((uint32_t)tp.ctrl << 16) | ((pm.z_write ? 1u : 0u) << 24);
            if (!is_clip_batch && tp.color < n_paints && r->paint_is_textured[tp.color]) pi.paint_ctrl |= PATH_TEXTURED;
            pi.clip_path_index = pm.clip_path_index;
            pi.pad = (uint32_t)i; // index in the command's arrays (initial backdrops are keyed by it)
            h_paths[k] = pi;
            h_seg_first[k] = seg_first;
            h_tile_off[k] = tile_off;
            h_col_off[k] = col_off;
            k++;
            seg_first += segments_of((uint32_t)i);
            tile_off += (uint32_t)((uint64_t)w * (uint64_t)h);
            col_off += (uint32_t)w;
        }
        if (bad) bad_paint_flag.store(true);
    });
    laps.lap("meta pass 2");
    if (bad_paint_flag.load()) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "paint id outside the uploaded texture metadata");
    uint32_t *h_seg_table = h_col_off + (P + 1);
    uint32_t *h_tile_table = h_seg_table + seg_table_n;
    uint32_t *h_col_table = h_tile_table + tile_table_n;
    // table[k] = largest p < kept with offsets[p] <= k << shift (offsets[0] == 0); chunks of the table
    // are independent: each starts from a binary search and sweeps forward.
    auto build_table = [kept](const uint32_t *offsets, size_t n, int shift, uint32_t *table) {
        parallel_ranges(n, 8192, [&](size_t begin, size_t end) {
            const uint64_t x0 = (uint64_t)begin << shift;
            uint32_t p = (uint32_t)(std::upper_bound(offsets, offsets + kept, x0,
                                                     [](uint64_t x, uint32_t o) { return x < (uint64_t)o; }) - offsets);
            p = p ? p - 1 : 0;
            for (size_t k = begin; k < end; k++) {
                const uint64_t x = (uint64_t)k << shift;
                while (p + 1 < kept && offsets[p + 1] <= x) p++;
                table[k] = p;
            }
        });
    };
    if (kept) {
        build_table(h_seg_first, seg_table_n, SEG_SHIFT, h_seg_table);
        build_table(h_tile_off, tile_table_n, TILE_SHIFT, h_tile_table);
        build_table(h_col_off, col_table_n, COL_SHIFT, h_col_table);
    }
    laps.lap("meta tables");
    r->batch_meta.ensure(meta_bytes + 64, 1.25);
    if (meta_bytes) {
        PF_CUDA_CHECK(cudaMemcpyAsync(r->batch_meta.ptr, hbase, meta_bytes, cudaMemcpyHostToDevice, st));
        if (!r->meta_copied) PF_CUDA_CHECK(cudaEventCreateWithFlags(&r->meta_copied, cudaEventDisableTiming));
        PF_CUDA_CHECK(cudaEventRecord(r->meta_copied, st));
    }
    r->stats.h2d_bytes += meta_bytes;
    laps.lap("meta enqueue copy");

    b = BatchDev{};
    b.points = segments.points.ptr;
    b.seg_indices = segments.indices.ptr;
    b.paths = reinterpret_cast<const PathInfo *>(r->batch_meta.ptr);
    b.path_seg_first = reinterpret_cast<const uint32_t *>(r->batch_meta.ptr + (size_t)P * sizeof(PathInfo));
    b.path_tile_offset = b.path_seg_first + (P + 1);
    b.path_col_offset = b.path_tile_offset + (P + 1);
    b.seg_index = CoarseIndex{b.path_col_offset + (P + 1), SEG_SHIFT};
    b.tile_index = CoarseIndex{b.seg_index.table + seg_table_n, TILE_SHIFT};
    b.col_index = CoarseIndex{b.tile_index.table + tile_table_n, COL_SHIFT};
    b.n_paths = kept;
    b.n_segments = n_segments;
    b.n_tiles = n_tiles;
    b.n_columns = n_cols;
    const PFTransform2F &t = info.transform;
    b.xf.m11 = t.matrix.m00, b.xf.m12 = t.matrix.m01, b.xf.m21 = t.matrix.m10, b.xf.m22 = t.matrix.m11;
    b.xf.tx = t.vector.x, b.xf.ty = t.vector.y;
    b.xf.identity = (b.xf.m11 == 1.0f && b.xf.m12 == 0.0f && b.xf.m21 == 0.0f && b.xf.m22 == 1.0f &&
                     b.xf.tx == 0.0f && b.xf.ty == 0.0f);
    // process_line_segment clips to scene.view_box() (tiler.rs:194): by default the framebuffer rect.
    b.view_box = r->has_view_box
                     ? r->view_box
                     : ViewBox{0.0f, 0.0f, (float)r->options.dest_size.x, (float)r->options.dest_size.y};
    b.fb = fb;

    // TransformCPUBinGPU-style non-zero initial backdrops (builder.rs:679-688).
    has_initial_backdrops = false;
    r->col_backdrop_init.ensure(n_cols + 1, 1.25);
    if (info.backdrops && info.backdrop_count) {
        std::vector<int32_t> init(n_cols, 0);
        std::vector<uint32_t> kept_index(P, 0xffffffffu);
        for (uint32_t k = 0; k < kept; k++) kept_index[h_paths[k].pad] = k;
        for (size_t i = 0; i < info.backdrop_count; i++) {
            const PFBackdropInfoD3D11 &bi = info.backdrops[i];
            if (bi.initial_backdrop == 0 || bi.path_index >= P || kept_index[bi.path_index] == 0xffffffffu) continue;
            const PathInfo &pi = h_paths[kept_index[bi.path_index]];
            if (bi.tile_x_offset < 0 || bi.tile_x_offset >= pi.max_x - pi.min_x) continue;
            init[pi.col_offset + (uint32_t)bi.tile_x_offset] = bi.initial_backdrop;
            has_initial_backdrops = true;
        }
        if (has_initial_backdrops) {
            PF_CUDA_CHECK(cudaMemcpyAsync(r->col_backdrop_init.ptr, init.data(), (size_t)n_cols * 4,
                                          cudaMemcpyHostToDevice, st));
            PF_CUDA_CHECK(cudaStreamSynchronize(st));
        }
    }
}

// Runs bound(device) -> dice -> bin -> propagate -> sort -> fill+tile for the cached batch.
// `sizing`: data-dependent counts (lines, fills, list entries) are read back as they become known
// and buffers grown to fit — three host syncs, like the reference's three read-backs
// (d3d11/renderer.rs:218-229,338-353,634-648). Otherwise the counts stay on the device, grids are
// sized from the previous frame's counts plus slack, and the only sync is the verification at the end.
// Returns false when a bound was exceeded (the caller re-runs in sizing mode).
bool finalize_batch(PFCudaRenderer *r);
void ensure_alpha_ids(PFCudaRenderer *r, const uint32_t *counts = nullptr, uint32_t alpha_base = 0);

// Lays the frame's zero-initialised arrays out in r->zeroed and clears all of them with one memset
// (eight separate clears cost more in launch gaps than in bandwidth on the small scenes).
void carve_zeroed(PFCudaRenderer *r, size_t n_paths, size_t n_tiles, size_t n_cols, size_t n_fb) {
    auto padded = [](size_t n) { return (n + 4) & ~(size_t)3; }; // >= n + 1, keeps every array 16-byte aligned
    const size_t total = 16 + padded(n_paths) + padded(n_tiles) + padded(n_cols) + 4 * padded(n_fb);
    r->zeroed.ensure(total, 1.25);
    uint32_t *p = r->zeroed.ptr;
    r->counters.ptr = p, p += 16;
    r->path_live.ptr = p, p += padded(n_paths);
    r->tile_word.ptr = p, p += padded(n_tiles);
    r->col_backdrop.ptr = reinterpret_cast<int32_t *>(p), p += padded(n_cols);
    r->z_buffer.ptr = reinterpret_cast<int32_t *>(p), p += padded(n_fb);
    r->fb_count.ptr = p, p += padded(n_fb);
    r->fb_cursor.ptr = p, p += padded(n_fb);
    r->fb_alpha.ptr = p;
    r->tile_queue.ensure(n_fb + 1);
    r->tile_queue_hdr.ensure(n_fb + 1);
    PF_CUDA_CHECK(cudaMemsetAsync(r->zeroed.ptr, 0, total * sizeof(uint32_t), r->stream));
}

bool run_pipeline(PFCudaRenderer *r, bool sizing, bool clip_pass = false) {
    cudaStream_t st = r->stream;
    BatchCache &c = r->cache;
    const BatchDev &b = c.batch;
    const FbRect fb = b.fb;
    const uint32_t n_segments = b.n_segments, n_tiles = b.n_tiles, n_cols = b.n_columns;
    int launches = 0;
    enum { C_LINES = 0, C_FILLS = 1, C_ENTRIES = 2, C_VISIBLE_FILLS = 5 };

    if (r->timing) {
        r->timer.create();
        PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[0], st));
    }
    // ---- bound (device part): clear the dense tile arrays. bound.cs.glsl:80-83 initialises
    // {next=-1, first_fill=-1, backdrop=0}; here a tile is one word (count | backdrop delta). The
    // counters, per-path flags, tile words, column backdrops, z-buffer and list counts share one
    // allocation and one memset.
    const int fb_w = fb.max_x - fb.min_x, fb_h = fb.max_y - fb.min_y;
    const uint32_t n_fb = (uint32_t)(fb_w * fb_h);
    carve_zeroed(r, b.n_paths, n_tiles, n_cols, n_fb);
    r->tile_fill_pos.ensure(n_tiles + 1, 1.25);
    r->tile_fb.ensure(n_tiles + 1, 1.25);
    r->fb_start.ensure(n_fb + 1);
    if (c.has_initial_backdrops)
        PF_CUDA_CHECK(cudaMemcpyAsync(r->col_backdrop.ptr, r->col_backdrop_init.ptr, (size_t)n_cols * 4,
                                      cudaMemcpyDeviceToDevice, st));
    if (r->debug_lists) {
        r->tile_first_fill.ensure(n_tiles + 1, 1.25);
        PF_CUDA_CHECK(cudaMemsetAsync(r->tile_first_fill.ptr, 0xff, (size_t)n_tiles * 4, st));
    }
    if (r->timing) PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[1], st));

    auto bound_of = [](uint32_t last, size_t capacity) -> uint32_t {
        uint64_t want = (uint64_t)last + last / 8 + 4096;
        return (uint32_t)(want < capacity ? want : capacity);
    };

    // ---- dice. Steady state: one pass that appends lines in arbitrary order (nothing but the parity
    // dumps depends on it). Sizing / parity dumps: count -> scan -> emit in segment order.
    const bool stream_dice = !sizing && !r->debug_lists;
    uint32_t line_bound;
    const uint32_t *n_lines_dev = r->counters.ptr + C_LINES;
    if (stream_dice) {
        line_bound = bound_of(c.n_lines, std::min(r->lines.capacity, r->line_path.capacity));
        launches += launch_dice_stream(b, r->lines.ptr, r->line_path.ptr, line_bound, r->counters.ptr + C_LINES, st);
    } else {
        r->seg_line_offset.ensure(n_segments + 1, 1.25);
        launches += launch_dice(false, b, r->seg_line_offset.ptr, nullptr, nullptr, nullptr, 0, st);
        launches += exclusive_scan(LoadU32{r->seg_line_offset.ptr}, r->seg_line_offset.ptr, n_segments,
                                   r->counters.ptr + C_LINES, r->scan_scratch, st);
        if (sizing) {
            line_bound = n_segments ? read_counter(r, C_LINES) : 0;
            r->lines.ensure(line_bound + 1, 1.25);
            r->line_path.ensure(line_bound + 1, 1.25);
            if (r->debug_lists) r->line_fill_offset.ensure(line_bound + 1, 1.25);
        } else {
            size_t cap = std::min(r->lines.capacity, r->line_path.capacity);
            if (r->debug_lists) cap = std::min(cap, r->line_fill_offset.capacity);
            line_bound = bound_of(c.n_lines, cap);
        }
        launches += launch_dice(true, b, nullptr, r->seg_line_offset.ptr, r->lines.ptr, r->line_path.ptr, line_bound, st);
    }
    if (r->timing) PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[2], st));

    // ---- bin, count pass: per-tile fill counts + backdrop deltas.
    BinArgs ba{};
    ba.lines = r->lines.ptr;
    ba.line_path = r->line_path.ptr;
    ba.n_lines = line_bound;
    ba.n_lines_dev = n_lines_dev;
    ba.tile_word = r->tile_word.ptr;
    ba.col_backdrop = r->col_backdrop.ptr;
    ba.line_fill_count = r->debug_lists ? r->line_fill_offset.ptr : nullptr;
    r->long_queue.ensure(LONG_QUEUE_CAPACITY);
    ba.long_queue = r->long_queue.ptr;
    ba.long_capacity = LONG_QUEUE_CAPACITY;
    ba.long_count = r->counters.ptr + 6;
    ba.long_cursor = r->counters.ptr + 7;
    launches += launch_bin(1 /* BIN_COUNT */, b, ba, st);
    ba.long_count = r->counters.ptr + 13; // the emit pass queues its long lines afresh
    ba.long_cursor = r->counters.ptr + 14;
    uint32_t emit_bound = 0;
    if (r->debug_lists) // emission-order offsets and the total fill count, for the parity dumps
        launches += exclusive_scan(LoadU32{r->line_fill_offset.ptr}, r->line_fill_offset.ptr, line_bound,
                                   r->counters.ptr + C_FILLS, r->scan_scratch, st, n_lines_dev);
    if (r->timing) PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[3], st));

    // ---- propagate: column backdrop prefix sums + occluder z-writes.
    // A batch with clipped paths resolves its tiles against the clip batch here (the four cases of
    // Tiler::prepare_tiles) and leaves a reference to the clip tile wherever a mask has to be combined.
    ClipDev clip_dev{};
    const bool use_clip = !clip_pass && c.has_clips;
    if (use_clip) {
        clip_dev.paths = reinterpret_cast<const PathInfo *>(r->clip.meta.ptr);
        clip_dev.n_paths = r->clip.n_paths;
        clip_dev.tile_word = r->clip.tile_word.ptr;
        clip_dev.tile_fill_end = r->clip.tile_fill_end.ptr;
        clip_dev.fills = r->clip.fills.ptr;
        r->tile_clip.ensure(n_tiles + 1, 1.25);
        if (r->debug_lists) r->tile_orig.ensure(n_tiles + 1, 1.25);
    }
    const bool clip_dumps = use_clip && r->debug_lists;
    launches += launch_propagate(b, r->tile_word.ptr, r->col_backdrop.ptr, r->z_buffer.ptr,
                                 use_clip ? &clip_dev : nullptr, use_clip ? r->tile_clip.ptr : nullptr,
                                 clip_dumps ? r->tile_orig.ptr : nullptr, st);
    if (r->timing) PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[4], st));

    // ---- sort: z-cull + per-framebuffer-tile runs (count -> scan -> append); the run itself is
    // sorted into draw order inside the fused kernel.
    // (the same kernel reserves the fill runs of the surviving tiles: occlusion culling before fill
    // emission; with the parity dumps on, every alpha tile keeps its fills so its mask can be read back)
    // Steady state: the survivors are also compacted into a list sized like the entries (same bound, same
    // overflow condition), so the entry pass does not have to visit every bbox tile.
    const bool compact_lists = !sizing && !r->debug_lists;
    uint32_t live_capacity = 0;
    if (compact_lists) {
        live_capacity = bound_of(c.n_entries, r->entries.capacity);
        r->live_tiles.ensure(live_capacity + 1, 1.25);
    }
    launches += launch_list_count(b, r->tile_word.ptr, r->z_buffer.ptr, r->tile_fb.ptr, r->fb_count.ptr,
                                  r->tile_fill_pos.ptr, r->counters.ptr + C_VISIBLE_FILLS, r->path_live.ptr,
                                  r->debug_lists, clip_dumps ? r->tile_orig.ptr : nullptr,
                                  compact_lists ? r->live_tiles.ptr : nullptr, live_capacity, r->counters.ptr + 15,
                                  r->fb_alpha.ptr, use_clip ? r->tile_clip.ptr : nullptr, st);
    launches += exclusive_scan(LoadU32{r->fb_count.ptr}, r->fb_start.ptr, n_fb, r->counters.ptr + C_ENTRIES,
                               r->scan_scratch, st);
    uint32_t entry_bound, fill_bound;
    if (sizing) {
        if (n_tiles) read_counter(r, 0); // one read-back of all totals
        entry_bound = n_tiles ? r->counters_host.ptr[C_ENTRIES] : 0;
        fill_bound = n_tiles ? r->counters_host.ptr[C_VISIBLE_FILLS] : 0;
        r->entries.ensure(entry_bound + 1, 1.25);
        r->fills.ensure(fill_bound + 1, 1.25);
        if (r->debug_lists) {
            emit_bound = n_tiles ? r->counters_host.ptr[C_FILLS] : 0;
            r->fills_emit.ensure(emit_bound + 1, 1.25);
        }
    } else {
        entry_bound = bound_of(c.n_entries, r->entries.capacity);
        fill_bound = bound_of(c.n_visible_fills, r->fills.capacity);
        if (r->debug_lists) emit_bound = bound_of(c.n_fills, r->fills_emit.capacity);
    }
    if (r->timing) PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[5], st));

    // ---- bin, emit pass: fills of surviving tiles into their tile-grouped runs.
    ba.tile_fb = r->tile_fb.ptr;
    ba.path_live = r->path_live.ptr;
    ba.tile_fill_pos = r->tile_fill_pos.ptr;
    ba.fills = r->fills.ptr;
    ba.fill_capacity = fill_bound;
    if (r->debug_lists) {
        ba.line_fill_offset = r->line_fill_offset.ptr;
        ba.tile_first_fill = r->tile_first_fill.ptr;
        ba.fills_emit = r->fills_emit.ptr;
        ba.emit_capacity = emit_bound;
        launches += launch_bin(2 /* BIN_EMIT */, b, ba, st);
    } else {
        launches += launch_bin(0 /* BIN_EMIT_LIVE */, b, ba, st);
    }
    if (clip_pass) {
        // The clip batch ends here: keep what the draw batch will read — the stage buffers themselves are
        // reused by the draw batch.
        r->clip.n_paths = b.n_paths;
        r->clip.meta.ensure((size_t)b.n_paths * sizeof(PathInfo) + 64);
        r->clip.tile_word.ensure(n_tiles + 1);
        r->clip.tile_fill_end.ensure(n_tiles + 1);
        r->clip.fills.ensure(fill_bound + 1);
        auto keep = [&](void *dst, const void *src, size_t bytes) {
            if (bytes) PF_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
        };
        keep(r->clip.meta.ptr, b.paths, (size_t)b.n_paths * sizeof(PathInfo));
        keep(r->clip.tile_word.ptr, r->tile_word.ptr, (size_t)n_tiles * 4);
        keep(r->clip.tile_fill_end.ptr, r->tile_fill_pos.ptr, (size_t)n_tiles * 4);
        keep(r->clip.fills.ptr, r->fills.ptr, (size_t)fill_bound * sizeof(PackedFill));
        r->clip.n_fills = r->clip.n_alpha = 0;
        if (r->debug_lists) {
            // Parity dumps: number the clip tiles' alpha ids now (clip paths precede draw paths in
            // SequentialExecutor order) and keep their fills as records.
            r->last_batch = b;
            r->last_fills = emit_bound;
            r->last_alpha_ids_valid = false;
            ensure_alpha_ids(r, r->tile_word.ptr, 0);
            r->clip.n_fills = emit_bound;
            r->clip.n_alpha = r->last_alpha_tiles;
            r->clip.tile_alpha_id.ensure(n_tiles + 1);
            r->clip.fill_records.ensure((size_t)emit_bound * sizeof(PFFill) + 16);
            keep(r->clip.tile_alpha_id.ptr, r->tile_alpha_id.ptr, (size_t)n_tiles * 4);
            launch_dump_fills(emit_bound, r->fills_emit.ptr, r->tile_alpha_id.ptr, r->clip.fill_records.ptr, st);
            r->last_alpha_ids_valid = false;
            r->stats.fill_count += emit_bound;
        }
        PF_CUDA_CHECK(cudaStreamSynchronize(st)); // (the clip pass always runs with exact sizes, synchronously)
        r->stats.host_sync_count++;
        r->stats.drawcall_count += (uint64_t)launches;
        r->clip.valid = true;
        return true;
    }
    if (use_clip) r->entry_clip.ensure(entry_bound + 1, 1.25);
    const OverflowGuard guard{r->counters.ptr, line_bound, entry_bound, fill_bound, r->counters.ptr + 12};
    launches += launch_list_emit(b, r->tile_fb.ptr, r->tile_word.ptr, r->tile_fill_pos.ptr, r->fb_start.ptr,
                                 r->fb_cursor.ptr, r->paints.ptr, r->entries.ptr, entry_bound, guard,
                                 use_clip ? &clip_dev : nullptr, use_clip ? r->tile_clip.ptr : nullptr,
                                 use_clip ? r->entry_clip.ptr : nullptr, compact_lists ? r->live_tiles.ptr : nullptr,
                                 live_capacity, r->counters.ptr + 15, st);
    if (r->timing) PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[6], st));

    // ---- fill + tile (fused).
    CompositeArgs ca{};
    ca.entries = r->entries.ptr;
    ca.entry_clip = use_clip ? r->entry_clip.ptr : nullptr;
    ca.clip_fills = use_clip ? r->clip.fills.ptr : nullptr;
    ca.fb_start = r->fb_start.ptr;
    ca.fb_count = r->fb_count.ptr;
    ca.fb_alpha = r->fb_alpha.ptr;
    ca.queue = r->tile_queue.ptr;
    ca.queue_hdr = r->tile_queue_hdr.ptr;
    ca.queue_count = r->counters.ptr + 10;
    ca.fills = r->fills.ptr;
    ca.area_lut = r->lut_tex;
    ca.fb = fb;
    ca.tile_y0 = c.strip_y0;
    ca.tile_y1 = c.strip_y1;
    const DrawTarget target = current_target(r);
    fill_destinations(r, target, ca);
    // A render target starts from transparent black (Renderer::clear_color_for_draw_operation, gpu/renderer.rs).
    ca.clear_color = target.is_main ? clear_color(r) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    ca.load_dest = *target.batches_drawn > 0;
    if (r->any_textured_paint) { // some paint samples the batch's colour texture or blends per pixel
        ca.paint_textures = r->paint_textures.ptr;
        ca.color_texture = c.color_texture;
        ca.gamma_lut = r->has_gamma_lut ? r->gamma_lut.ptr : nullptr;
    }
    ca.work_counter = r->counters.ptr + 12;
    wait_for_gather_before_compositing(r, st, r->gather_last_was_tiles);
    if (target.is_main) {
        // Tile mode: the (first) destination batch of the frame also writes the compact export of this strip.
        r->main_batches_this_frame++;
        r->frame_exported = false;
        if (r->gather_comm && r->gather_mode == PF_CUDA_GATHER_MODE_TILES && r->main_batches_this_frame == 1 && !ca.load_dest &&
            r->peer_export[r->gather_rank]) {
            const int parity = (int)(r->gather_frame_serial & 1);
            if (r->push_recorded[parity]) PF_CUDA_CHECK(cudaStreamWaitEvent(st, r->push_done[parity], 0));
            uint8_t *buffer = r->export_region.ptr + ((size_t)r->gather_rank * 2 + parity) * r->export_buffer_bytes;
            ca.export_alpha_count = reinterpret_cast<uint32_t *>(buffer);
            ca.queue = reinterpret_cast<uint32_t *>(buffer + r->export_queue_off);
            ca.export_solid_color = reinterpret_cast<uint32_t *>(buffer + r->export_color_off);
            ca.export_solid_mask = reinterpret_cast<uint32_t *>(buffer + r->export_mask_off);
            ca.export_blocks = buffer + r->export_blocks_off;
            r->frame_exported = true;
        }
    }
    launches += launch_composite(ca, st);
    if (r->timing) PF_CUDA_CHECK(cudaEventRecord(r->timer.ev[7], st));

    // ---- verification: one read-back of the totals at the end of the batch. With deferred
    // verification the host does not wait for it here: the totals are checked the next time the
    // renderer is used (verify_pending), so consecutive frames are enqueued back to back.
    r->counters_host.ensure(16);
    PF_CUDA_CHECK(cudaMemcpyAsync(r->counters_host.ptr, r->counters.ptr, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PendingVerify &pv = r->pending;
    pv.line_bound = line_bound, pv.fill_bound = fill_bound, pv.entry_bound = entry_bound, pv.emit_bound = emit_bound;
    pv.launches = launches;
    pv.batches_drawn_before = *target.batches_drawn;
    if (r->deferred_verify && !sizing) {
        if (!r->verify_event) PF_CUDA_CHECK(cudaEventCreateWithFlags(&r->verify_event, cudaEventDisableTiming));
        PF_CUDA_CHECK(cudaEventRecord(r->verify_event, st));
        pv.active = true;
        return true; // optimistic; verify_pending() repairs the frame if a bound was exceeded
    }
    PF_CUDA_CHECK(cudaStreamSynchronize(st));
    r->borrowed_copies_pending = false;
    r->stats.host_sync_count++;
    return finalize_batch(r);
}

// Reads the totals of the batch whose counters have arrived in pinned memory, updates caches / stats /
// stage times, and reports whether every stage stayed inside its bound.
bool finalize_batch(PFCudaRenderer *r) {
    enum { C_LINES = 0, C_FILLS = 1, C_ENTRIES = 2, C_VISIBLE_FILLS = 5 };
    BatchCache &c = r->cache;
    const BatchDev &b = c.batch;
    const PendingVerify &pv = r->pending;
    const uint32_t n_lines = r->counters_host.ptr[C_LINES], n_fills = r->counters_host.ptr[C_FILLS],
                   n_entries = r->counters_host.ptr[C_ENTRIES], n_visible = r->counters_host.ptr[C_VISIBLE_FILLS];
    r->stats.drawcall_count += (uint64_t)pv.launches;
    c.n_lines = n_lines;
    c.n_fills = n_fills;
    c.n_entries = n_entries;
    c.n_visible_fills = n_visible;
    if (n_lines > pv.line_bound || n_visible > pv.fill_bound || n_entries > pv.entry_bound ||
        (r->debug_lists && n_fills > pv.emit_bound)) {
        c.counts_valid = false;
        return false;
    }
    c.counts_valid = true;

    r->last_batch = b;
    r->last_lines = n_lines;
    r->last_fills = n_fills;
    r->last_entries = n_entries;
    r->last_alpha_ids_valid = false;
    r->last_fb = b.fb;
    r->stats.path_count += b.n_paths;
    r->stats.fill_count += n_fills;
    r->stats.total_tile_count += b.n_tiles;
    r->stats.input_segment_count += b.n_segments;
    r->stats.line_segment_count += n_lines;
    r->stats.tile_list_entry_count += n_entries;
    r->stats.visible_fill_count += n_visible;
    r->stats.column_count += b.n_columns;

    if (r->timing) {
        float ms[7];
        for (int i = 0; i < 7; i++) PF_CUDA_CHECK(cudaEventElapsedTime(&ms[i], r->timer.ev[i], r->timer.ev[i + 1]));
        r->times.bound_ms += ms[0];
        r->times.dice_ms += ms[1];
        r->times.bin_ms += ms[2] + ms[5]; // count pass + emit pass (the emit pass runs after the z-cull)
        r->times.propagate_ms += ms[3];
        r->times.sort_ms += ms[4];
        r->times.fill_tile_ms += ms[6];
        float total;
        PF_CUDA_CHECK(cudaEventElapsedTime(&total, r->timer.ev[0], r->timer.ev[7]));
        r->times.total_ms += total;
        // since timing was switched on (not reset per frame: frames verified late still count)
        r->times_total.bound_ms += ms[0];
        r->times_total.dice_ms += ms[1];
        r->times_total.bin_ms += ms[2] + ms[5];
        r->times_total.propagate_ms += ms[3];
        r->times_total.sort_ms += ms[4];
        r->times_total.fill_tile_ms += ms[6];
        r->times_total.total_ms += total;
        r->times_batches++;
    }
    return true;
}

// Deferred verification: waits for the totals of the last batch (normally long finished), and if a
// stage overflowed its bound re-renders that batch with exact sizing.
void verify_pending(PFCudaRenderer *r) {
    if (!r->pending.active) return;
    r->pending.active = false;
    PF_CUDA_CHECK(cudaEventSynchronize(r->verify_event));
    r->borrowed_copies_pending = false; // the event was recorded after every copy of that frame
    r->stats.host_sync_count++;
    if (finalize_batch(r)) return;
    int &drawn = *current_target(r).batches_drawn; // (verification is never deferred across a target change)
    const int drawn_now = drawn;
    drawn = r->pending.batches_drawn_before; // same load action as the failed attempt
    const bool ok = run_pipeline(r, true);
    drawn = drawn_now;
    r->stats.reruns++;
    if (!ok) throw Error(PF_CUDA_ERROR_CUDA, "stage buffer overflow after exact sizing (internal error)");
}

// One DrawTilesD3D11 batch: prepare_tiles + draw_tiles (d3d11/renderer.rs:414-424).
void verify_pending(PFCudaRenderer *r);

// PrepareClipTilesD3D11: dice, bin and propagate the clip paths and keep their tiles and fills for the
// draw batch that references them (d3d11/renderer.rs prepare_tiles on a clip batch).
// One level of clipping: a clip path that is itself clipped is refused.
void prepare_clip_batch(PFCudaRenderer *r, const PFTileBatchDataD3D11 &batch) {
    verify_pending(r);
    if (!r->has_scene) throw Error(PF_CUDA_ERROR_PROTOCOL, "PrepareClipTilesD3D11 before UploadSceneD3D11");
    if (batch.path_source != PF_PATH_SOURCE_CLIP)
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "clip batch with a draw path source");
    if (batch.has_clipped_path_info && batch.clipped_path_info.clipped_path_count > 0)
        throw Error(PF_CUDA_ERROR_UNSUPPORTED, "nested clip paths are not implemented");
    if (r->clip.valid) throw Error(PF_CUDA_ERROR_UNSUPPORTED, "more than one clip batch per frame (nested clip levels)");
    if (!r->target_stack.empty())
        throw Error(PF_CUDA_ERROR_UNSUPPORTED, "clip batches inside a render target are not implemented");
    const FbRect fb = framebuffer_tile_rect(r);
    const int32_t strip_y0 = r->strip_y1 > r->strip_y0 ? r->strip_y0 : fb.min_y;
    const int32_t strip_y1 = r->strip_y1 > r->strip_y0 ? r->strip_y1 : fb.max_y;
    BatchCache &c = r->cache;
    c.valid = false; // the stage buffers and the batch slot are borrowed; the draw batch re-uploads
    c.has_clips = false;
    c.color_texture = ColorTexture{nullptr, 0, 0, 0, 0, 0};
    upload_batch_metadata(r, batch, fb, strip_y0, strip_y1, c.batch, c.has_initial_backdrops, r->clip_segments, true);
    c.strip_y0 = strip_y0;
    c.strip_y1 = strip_y1;
    run_pipeline(r, true, true);
}

// DrawTileBatchD3D11.color_texture -> the page it names (Renderer::draw_tiles binds it as uColorTexture0,
// renderer/src/gpu/d3d11/renderer.rs:733-741).
ColorTexture resolve_color_texture(PFCudaRenderer *r, bool has_color_texture, const PFTileBatchTexture &texture) {
    if (!has_color_texture) return ColorTexture{nullptr, 0, 0, 0, 0, 0};
    if (texture.page >= r->pages.size() || !r->pages[texture.page])
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "draw batch names a texture page that was never allocated");
    // (TileBatchTexture::composite_op travels with the batch but no renderer of the reference reads it: how the
    // texture combines with the base colour is the texture metadata entry's color_0_combine_mode, paint.rs:649-653)
    if (texture.composite_op > PF_PAINT_COMPOSITE_OP_DEST_IN)
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "unknown PaintCompositeOp");
    const PFCudaRenderer::TexturePage &page = *r->pages[texture.page];
    if (!r->target_stack.empty() && r->render_targets[r->target_stack.back()].page == texture.page)
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "draw batch samples the render target it draws to");
    return ColorTexture{page.pixels.ptr, (size_t)page.width * 4, page.width, page.height, page.is_render_target ? 1 : 0,
                        (uint32_t)texture.sampling_flags};
}

void draw_tile_batch(PFCudaRenderer *r, const PFTileBatchDataD3D11 &batch, bool has_color_texture,
                     const PFTileBatchTexture &texture) {
    verify_pending(r);
    if (!r->has_scene) throw Error(PF_CUDA_ERROR_PROTOCOL, "DrawTilesD3D11 before UploadSceneD3D11");
    if (batch.path_source != PF_PATH_SOURCE_DRAW)
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "draw batch with a clip path source");
    const bool has_clips = batch.has_clipped_path_info && batch.clipped_path_info.clipped_path_count > 0;
    if (has_clips && !r->clip.valid)
        throw Error(PF_CUDA_ERROR_PROTOCOL, "draw batch with clipped paths before PrepareClipTilesD3D11");
    const DrawTarget target = current_target(r);
    const ColorTexture color_texture = resolve_color_texture(r, has_color_texture, texture);
    if (has_clips && !target.is_main)
        throw Error(PF_CUDA_ERROR_UNSUPPORTED, "clipped paths inside a render target are not implemented");
    const FbRect fb = target_tile_rect(target);
    // Strips partition the destination image; a render target is needed whole by whichever rank samples it.
    const bool strip = target.is_main && r->strip_y1 > r->strip_y0;
    const int32_t strip_y0 = strip ? r->strip_y0 : fb.min_y;
    const int32_t strip_y1 = strip ? r->strip_y1 : fb.max_y;
    const ViewBox vb = r->has_view_box
                           ? r->view_box
                           : ViewBox{0.0f, 0.0f, (float)r->options.dest_size.x, (float)r->options.dest_size.y};

    // Batch cache: a non-zero content_key promises that the metadata arrays are identical to the
    // previous batch with the same key (the "scene not dirty" case extended to the per-frame batch
    // data, which the reference rebuilds and re-uploads every frame, builder.rs:653-759).
    BatchCache &c = r->cache;
    const bool hit = c.valid && batch.content_key != 0 && batch.content_key == c.key &&
                     c.scene_generation == r->scene_generation && c.paint_generation == r->paint_generation &&
                     c.strip_y0 == strip_y0 && c.strip_y1 == strip_y1 &&
                     memcmp(&c.batch.fb, &fb, sizeof(fb)) == 0 && memcmp(&c.batch.view_box, &vb, sizeof(vb)) == 0 &&
                     c.command_paths == batch.path_count && c.command_segments == batch.segment_count;
    if (!hit) {
        // A batch of the same shape as the previous one (same path / segment counts, same strip)
        // is most likely the same scene re-sent: keep the previous totals as bounds so the frame
        // runs without count read-backs; the end-of-batch verification falls back to exact sizing.
        const bool same_shape = c.valid && c.counts_valid && c.command_paths == batch.path_count &&
                                c.command_segments == batch.segment_count && c.strip_y0 == strip_y0 &&
                                c.strip_y1 == strip_y1 && memcmp(&c.batch.fb, &fb, sizeof(fb)) == 0;
        c.valid = false;
        upload_batch_metadata(r, batch, fb, strip_y0, strip_y1, c.batch, c.has_initial_backdrops, r->draw_segments, false);
        c.key = batch.content_key;
        c.command_paths = batch.path_count;
        c.command_segments = batch.segment_count;
        c.scene_generation = r->scene_generation;
        c.paint_generation = r->paint_generation;
        c.strip_y0 = strip_y0;
        c.strip_y1 = strip_y1;
        c.counts_valid = same_shape && !has_clips; // batches with clips are sized exactly every frame
        c.has_clips = has_clips;
        c.valid = true;
    } else {
        r->stats.batch_cache_hits++;
    }
    c.color_texture = color_texture;
    bool ok = run_pipeline(r, !c.counts_valid || r->always_size);
    if (!ok) {
        ok = run_pipeline(r, true);
        r->stats.reruns++;
        if (!ok) throw Error(PF_CUDA_ERROR_CUDA, "stage buffer overflow after exact sizing (internal error)");
    }
    (*target.batches_drawn)++;
}

// ---- texture pages and render targets (Renderer::allocate_pattern_texture_page / upload_texel_data /
// declare_render_target / push_render_target / pop_render_target, renderer/src/gpu/renderer.rs:462-520).
void allocate_texture_page(PFCudaRenderer *r, uint32_t page_id, PFVector2I size) {
    if (page_id >= 4096 || size.x <= 0 || size.y <= 0 || (int64_t)size.x * size.y > (1ll << 30))
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "bad texture page id / size");
    if (r->pages.size() <= page_id) r->pages.resize(page_id + 1);
    if (!r->pages[page_id]) r->pages[page_id].reset(new PFCudaRenderer::TexturePage());
    PFCudaRenderer::TexturePage &page = *r->pages[page_id];
    if (page.width == size.x && page.height == size.y) return; // the reference re-allocates only on a size change
    verify_pending(r);
    PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    page.pixels.bytes_allocated = &r->bytes_allocated;
    page.pixels.ensure((size_t)size.x * size.y * 4);
    page.width = size.x, page.height = size.y;
    PF_CUDA_CHECK(cudaMemsetAsync(page.pixels.ptr, 0, (size_t)size.x * size.y * 4, r->stream));
}

const PFCudaRenderer::TexturePage &page_of(PFCudaRenderer *r, const PFTextureLocation &loc, const char *what) {
    if (loc.page >= r->pages.size() || !r->pages[loc.page])
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, std::string(what) + ": texture page was never allocated");
    const PFCudaRenderer::TexturePage &page = *r->pages[loc.page];
    if (loc.rect.origin.x < 0 || loc.rect.origin.y < 0 || loc.rect.lower_right.x > page.width ||
        loc.rect.lower_right.y > page.height || loc.rect.lower_right.x <= loc.rect.origin.x ||
        loc.rect.lower_right.y <= loc.rect.origin.y)
        throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, std::string(what) + ": rectangle outside its texture page");
    return page;
}

void upload_texel_data(PFCudaRenderer *r, const PFColorU *texels, size_t count, const PFTextureLocation &loc) {
    const PFCudaRenderer::TexturePage &page = page_of(r, loc, "UploadTexelData");
    const size_t w = (size_t)(loc.rect.lower_right.x - loc.rect.origin.x), h = (size_t)(loc.rect.lower_right.y - loc.rect.origin.y);
    if (!texels || count != w * h) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "UploadTexelData: texel count does not match the rectangle");
    uint8_t *dst = page.pixels.ptr + ((size_t)loc.rect.origin.y * page.width + (size_t)loc.rect.origin.x) * 4;
    PF_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)page.width * 4, texels, w * 4, w * 4, h, cudaMemcpyHostToDevice, r->stream));
    PF_CUDA_CHECK(cudaStreamSynchronize(r->stream)); // the payload is borrowed for the call
    r->stats.h2d_bytes += w * h * 4;
}

void declare_render_target(PFCudaRenderer *r, uint32_t id, const PFTextureLocation &loc) {
    if (id >= 4096) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "bad render target id");
    page_of(r, loc, "DeclareRenderTarget");
    if (r->render_targets.size() <= id) r->render_targets.resize(id + 1);
    PFCudaRenderer::RenderTargetSlot &slot = r->render_targets[id];
    slot.declared = true;
    slot.page = loc.page;
    slot.rect = loc.rect;
    slot.batches_drawn = 0;
    r->pages[loc.page]->is_render_target = true;
}

// Alpha tile ids in SequentialExecutor order for the last batch (needs debug lists).
// `counts`: where the tiles' fill counts are read from (NULL: the tile words, or with clipped paths the
// counts before the clip was applied); `alpha_base`: ids already taken by the clip batch.
void ensure_alpha_ids(PFCudaRenderer *r, const uint32_t *counts, uint32_t alpha_base) {
    if (r->last_alpha_ids_valid) return;
    if (!counts) {
        const bool clipped = r->cache.has_clips && r->clip.valid;
        counts = clipped ? r->tile_orig.ptr : r->tile_word.ptr;
        alpha_base = clipped ? r->clip.n_alpha : 0;
    }
    if (!r->debug_lists) throw Error(PF_CUDA_ERROR_PROTOCOL, "enable debug lists before rendering to read alpha tile ids");
    cudaStream_t st = r->stream;
    const uint32_t n_tiles = r->last_batch.n_tiles, n_fills = r->last_fills;
    r->fill_is_first.ensure(n_fills + 1, 1.25);
    r->fill_first_scan.ensure(n_fills + 1, 1.25);
    r->tile_alpha_id.ensure(n_tiles + 1, 1.25);
    PF_CUDA_CHECK(cudaMemsetAsync(r->fill_is_first.ptr, 0, n_fills, st));
    launch_alpha_flags(n_tiles, counts, r->tile_first_fill.ptr, r->fill_is_first.ptr, st);
    exclusive_scan(LoadU8{r->fill_is_first.ptr}, r->fill_first_scan.ptr, n_fills, r->counters.ptr + 3,
                   r->scan_scratch, st);
    launch_alpha_assign(n_tiles, counts, r->tile_first_fill.ptr, r->fill_first_scan.ptr, r->tile_alpha_id.ptr,
                        alpha_base, st);
    r->last_alpha_tiles = n_fills ? read_counter(r, 3) : 0;
    r->last_alpha_ids_valid = true;
}

void close_peers(PFCudaRenderer *r) {
    for (int i = 0; i < r->n_peers; i++)
        if (r->peer_base[i]) cudaIpcCloseMemHandle(r->peer_base[i]);
    r->n_peers = 0;
}


// ---- NCCL, resolved at run time: the library must load (and every single-GPU path work) where libnccl is absent.
struct Nccl {
    typedef int (*GetUniqueIdFn)(void *);
    typedef int (*CommInitRankFn)(void **, int, PFCudaGatherId, int); // ncclUniqueId is 128 bytes passed by value
    typedef int (*CommDestroyFn)(void *);
    typedef int (*AllGatherFn)(const void *, void *, size_t, int, void *, cudaStream_t);
    typedef int (*BroadcastFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    typedef int (*GroupFn)();
    typedef const char *(*ErrorStringFn)(int);
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllGatherFn all_gather = nullptr;
    BroadcastFn broadcast = nullptr;
    AllReduceFn all_reduce = nullptr;
    GroupFn group_start = nullptr, group_end = nullptr;
    ErrorStringFn error_string = nullptr;
    bool ok = false;
    static constexpr int UINT8 = 1, UINT32 = 3, SUM = 0; // ncclUint8, ncclUint32, ncclSum

    static const Nccl &get() {
        static const Nccl instance = load();
        return instance;
    }
    static Nccl load() {
        Nccl n;
        // A process that already has NCCL (PyTorch bundles its own copy) must share it: two copies of the
        // library would each open the devices' NVLink resources.
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return n;
        n.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
        n.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
        n.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
        n.all_gather = (AllGatherFn)dlsym(h, "ncclAllGather");
        n.broadcast = (BroadcastFn)dlsym(h, "ncclBroadcast");
        n.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
        n.group_start = (GroupFn)dlsym(h, "ncclGroupStart");
        n.group_end = (GroupFn)dlsym(h, "ncclGroupEnd");
        n.error_string = (ErrorStringFn)dlsym(h, "ncclGetErrorString");
        n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_gather && n.broadcast && n.all_reduce && n.group_start &&
               n.group_end && n.error_string;
        return n;
    }
    void check(int result, const char *what) const {
        if (result != 0) throw Error(PF_CUDA_ERROR_CUDA, std::string(what) + " failed: " + error_string(result));
    }
};

const Nccl &nccl_or_throw() {
    const Nccl &n = Nccl::get();
    if (!n.ok) throw Error(PF_CUDA_ERROR_UNSUPPORTED, "libnccl.so.2 not found: frame assembly across GPUs is unavailable");
    return n;
}

void strip_of_rank(int32_t rows, int32_t rank, int32_t world, int32_t &y0, int32_t &y1) {
    y0 = (int32_t)((int64_t)rows * rank / world);
    y1 = (int32_t)((int64_t)rows * (rank + 1) / world);
}

void gather_destroy(PFCudaRenderer *r) {
    if (!r->gather_comm) return;
    cudaStreamSynchronize(r->gather_stream);
    Nccl::get().comm_destroy(r->gather_comm);
    r->gather_comm = nullptr;
    r->gather_in_flight = false;
    cudaEventDestroy(r->gather_ready);
    cudaEventDestroy(r->gather_done);
    if (r->gather_barrier_done) cudaEventDestroy(r->gather_barrier_done);
    r->gather_barrier_done = nullptr;
    for (int q = 0; q < 2; q++) {
        if (r->push_done[q]) cudaEventDestroy(r->push_done[q]);
        r->push_done[q] = nullptr;
        r->push_recorded[q] = false;
    }
    for (int g = 0; g < 8; g++) {
        if (r->peer_export_base[g]) cudaIpcCloseMemHandle(r->peer_export_base[g]);
        r->peer_export_base[g] = nullptr;
        r->peer_export[g] = nullptr;
    }
    r->export_region.release();
    if (r->barrier_word) cudaFree(r->barrier_word);
    r->barrier_word = nullptr;
    r->gather_last_was_tiles = false;
    cudaStreamDestroy(r->gather_stream);
    r->gather_stream = nullptr;
    r->gather_ready = r->gather_done = nullptr;
    r->gather_world = 0;
    r->strip_y0 = r->strip_y1 = 0;
}

template <typename F>
PFCudaStatus guarded(PFCudaRenderer *r, F &&f) {
    try {
        if (!r) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null renderer");
        r->bind_device();
        f();
        return PF_CUDA_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.status;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return PF_CUDA_ERROR_CUDA;
    }
}

template <typename F>
int64_t guarded_count(PFCudaRenderer *r, F &&f) {
    int64_t n = -1;
    PFCudaStatus s = guarded(r, [&]() { n = f(); });
    return s == PF_CUDA_OK ? n : -(int64_t)s;
}

} // namespace

extern "C" {

const char *PFCudaGetLastError(void) { return g_last_error.c_str(); }

PFCudaDeviceRef PFCudaDeviceCreate(int32_t ordinal) {
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        set_last_error(std::string("no CUDA device: ") + cudaGetErrorString(err));
        return nullptr;
    }
    if (ordinal < 0 || ordinal >= count) {
        set_last_error("CUDA device ordinal out of range");
        return nullptr;
    }
    return new PFCudaDevice{ordinal};
}

void PFCudaDeviceDestroy(PFCudaDeviceRef device) { delete device; }

uint8_t PFCudaDeviceGetFeatureLevel(PFCudaDeviceRef) { return PF_RENDERER_LEVEL_D3D11; }

PFCudaRendererRef PFCudaRendererCreate(PFCudaDeviceRef device, const uint8_t *area_lut_rgba8, const uint8_t *gamma_lut_l8,
                                       const PFRendererMode *mode, const PFCudaRendererOptions *options) {
    if (!device || !area_lut_rgba8 || !mode || !options) {
        set_last_error("PFCudaRendererCreate: null argument");
        return nullptr;
    }
    if (mode->level != PF_RENDERER_LEVEL_D3D11) {
        set_last_error("the CUDA backend implements RendererLevel::D3D11 only");
        return nullptr;
    }
    std::unique_ptr<PFCudaRenderer> r(new PFCudaRenderer());
    r->ordinal = device->ordinal;
    delete device; // ownership taken, as PFGLRendererCreate does (c/src/lib.rs:601-615)
    try {
        r->bind_device();
        setup_tracking(r.get());
        PF_CUDA_CHECK(cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking));
        r->stream = r->own_stream;
        r->options = *options;
        allocate_dest(r.get());
        // Area LUT: 256x256 RGBA8, LINEAR + CLAMP_TO_EDGE (gl/src/lib.rs:375,674-705).
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
        PF_CUDA_CHECK(cudaMallocArray(&r->lut_array, &desc, 256, 256));
        PF_CUDA_CHECK(cudaMemcpy2DToArray(r->lut_array, 0, 0, area_lut_rgba8, 256 * 4, 256 * 4, 256, cudaMemcpyHostToDevice));
        cudaResourceDesc res{};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = r->lut_array;
        cudaTextureDesc tex{};
        tex.addressMode[0] = cudaAddressModeClamp;
        tex.addressMode[1] = cudaAddressModeClamp;
        tex.filterMode = cudaFilterModeLinear;
        tex.readMode = cudaReadModeNormalizedFloat;
        tex.normalizedCoords = 1;
        PF_CUDA_CHECK(cudaCreateTextureObject(&r->lut_tex, &res, &tex, nullptr));
        if (gamma_lut_l8) { // textures/gamma-lut.png: 256 x 8, one byte per texel (gpu/renderer.rs:215-222)
            r->gamma_lut.ensure(256 * 8);
            PF_CUDA_CHECK(cudaMemcpy(r->gamma_lut.ptr, gamma_lut_l8, 256 * 8, cudaMemcpyHostToDevice));
            r->has_gamma_lut = true;
        }
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return nullptr;
    }
    return r.release();
}

void PFCudaRendererDestroy(PFCudaRendererRef r) {
    if (!r) return;
    cudaSetDevice(r->ordinal);
    cudaStreamSynchronize(r->stream);
    close_peers(r);
    gather_destroy(r);
    if (r->verify_event) cudaEventDestroy(r->verify_event);
    if (r->lut_tex) cudaDestroyTextureObject(r->lut_tex);
    if (r->lut_array) cudaFreeArray(r->lut_array);
    if (r->meta_copied) cudaEventDestroy(r->meta_copied);
    r->timer.destroy();
    cudaStream_t own = r->own_stream;
    delete r;
    if (own) cudaStreamDestroy(own);
}

PFCudaStatus PFCudaRendererSetOptions(PFCudaRendererRef r, const PFCudaRendererOptions *options) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (!options) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null options");
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        // A caller-provided (or peer) destination was validated for the size it was installed with: a frame
        // that no longer fits its rows or its height must not be composited into it.
        if ((r->dest_external || r->n_peers > 0) &&
            ((size_t)options->dest_size.x * 4 > r->dest_pitch || options->dest_size.y > r->dest_external_rows))
            throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT,
                        "dest_size exceeds the installed external destination (reset it with SetDestDevicePointer first)");
        r->options = *options;
        allocate_dest(r);
    });
}

PFCudaStatus PFCudaRendererBeginScene(PFCudaRendererRef r) {
    return guarded(r, [&]() {
        if (r->in_scene) throw Error(PF_CUDA_ERROR_PROTOCOL, "begin_scene called twice");
        verify_pending(r); // the previous frame's totals (deferred verification)
        r->in_scene = true;
        r->clip.valid = false;
        r->batches_drawn = 0;
        r->target_stack.clear();
        for (auto &slot : r->render_targets) slot.batches_drawn = 0;
        r->main_batches_this_frame = 0;
        r->frame_exported = false;
        r->stats = PFCudaRenderStats{};
        r->times = PFCudaRenderTime{};
    });
}

PFCudaStatus PFCudaRendererRenderCommand(PFCudaRendererRef r, const PFRenderCommand *cmd) {
    return guarded(r, [&]() {
        if (!cmd) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null command");
        if (!r->in_scene) throw Error(PF_CUDA_ERROR_PROTOCOL, "render_command outside begin_scene/end_scene");
        switch (cmd->kind) {
        case PF_RENDER_COMMAND_START:
            break;
        case PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA:
            // A non-zero content_key equal to the last upload's means "same table": skip it.
            if (cmd->u.upload_texture_metadata.content_key == 0 ||
                cmd->u.upload_texture_metadata.content_key != r->paint_key) {
                upload_texture_metadata(r, cmd->u.upload_texture_metadata.entries,
                                        cmd->u.upload_texture_metadata.entry_count);
                r->paint_key = cmd->u.upload_texture_metadata.content_key;
                r->paint_generation++;
            }
            break;
        case PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11:
            upload_segments(r, r->draw_segments, cmd->u.upload_scene_d3d11.draw_segments,
                            cmd->u.upload_scene_d3d11.payload_persists != 0);
            upload_segments(r, r->clip_segments, cmd->u.upload_scene_d3d11.clip_segments,
                            cmd->u.upload_scene_d3d11.payload_persists != 0);
            r->has_scene = true;
            r->scene_generation++;
            r->uploaded_scene_this_frame = true;
            break;
        case PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11:
            if (cmd->u.prepare_clip_tiles_d3d11.batch.path_count > 0)
                prepare_clip_batch(r, cmd->u.prepare_clip_tiles_d3d11.batch);
            break;
        case PF_RENDER_COMMAND_DRAW_TILES_D3D11:
            draw_tile_batch(r, cmd->u.draw_tiles_d3d11.tile_batch_data, cmd->u.draw_tiles_d3d11.has_color_texture != 0,
                            cmd->u.draw_tiles_d3d11.color_texture);
            break;
        case PF_RENDER_COMMAND_ALLOCATE_TEXTURE_PAGE:
            allocate_texture_page(r, cmd->u.allocate_texture_page.page_id, cmd->u.allocate_texture_page.size);
            break;
        case PF_RENDER_COMMAND_UPLOAD_TEXEL_DATA:
            upload_texel_data(r, cmd->u.upload_texel_data.texels, cmd->u.upload_texel_data.texel_count,
                              cmd->u.upload_texel_data.location);
            break;
        case PF_RENDER_COMMAND_DECLARE_RENDER_TARGET:
            declare_render_target(r, cmd->u.declare_render_target.render_target_id, cmd->u.declare_render_target.location);
            break;
        case PF_RENDER_COMMAND_PUSH_RENDER_TARGET: {
            const uint32_t id = cmd->u.push_render_target.render_target_id;
            if (id >= r->render_targets.size() || !r->render_targets[id].declared)
                throw Error(PF_CUDA_ERROR_PROTOCOL, "PushRenderTarget before DeclareRenderTarget");
            verify_pending(r); // a deferred verification belongs to the target it drew to
            r->target_stack.push_back(id);
            break;
        }
        case PF_RENDER_COMMAND_POP_RENDER_TARGET:
            if (r->target_stack.empty()) throw Error(PF_CUDA_ERROR_PROTOCOL, "PopRenderTarget without PushRenderTarget");
            verify_pending(r);
            r->target_stack.pop_back();
            break;
        case PF_RENDER_COMMAND_FINISH:
            r->stats.cpu_build_time_ns = cmd->u.finish.cpu_build_time_ns;
            break;
        case PF_RENDER_COMMAND_ADD_FILLS_D3D9:
        case PF_RENDER_COMMAND_FLUSH_FILLS_D3D9:
        case PF_RENDER_COMMAND_DRAW_TILES_D3D9:
            // Renderer::require_d3d11 (gpu/renderer.rs:1349-1360) panics here.
            throw Error(PF_CUDA_ERROR_WRONG_LEVEL, "D3D9-level command sent to the D3D11-level CUDA renderer");
        default:
            throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "unknown render command kind");
        }
    });
}

PFCudaStatus PFCudaRendererEndScene(PFCudaRendererRef r) {
    return guarded(r, [&]() {
        if (!r->in_scene) throw Error(PF_CUDA_ERROR_PROTOCOL, "end_scene without begin_scene");
        r->in_scene = false;
        if (!r->target_stack.empty()) {
            r->target_stack.clear();
            throw Error(PF_CUDA_ERROR_PROTOCOL, "end_scene with a render target still pushed");
        }
        if (r->batches_drawn == 0) {
            // Nothing drawn: the frame is the clear colour (tile.cs.glsl LOAD_ACTION_CLEAR).
            const FbRect fb = framebuffer_tile_rect(r);
            const uint32_t n_fb = (uint32_t)((fb.max_x - fb.min_x) * (fb.max_y - fb.min_y));
            r->fb_start.ensure(n_fb + 1);
            carve_zeroed(r, 0, 0, 0, n_fb); // every list empty, tile counter zero
            r->last_fb = fb;                // (the z-buffer of an empty frame is all zero over the same rect)
            r->last_batch = BatchDev{};
            r->last_lines = r->last_fills = r->last_entries = 0;
            r->last_alpha_ids_valid = false;
            CompositeArgs ca{};
            ca.fb_start = r->fb_start.ptr;
            ca.fb_count = r->fb_count.ptr;
            ca.fb_alpha = r->fb_alpha.ptr;
            ca.queue = r->tile_queue.ptr;
            ca.queue_hdr = r->tile_queue_hdr.ptr;
            ca.queue_count = r->counters.ptr + 10;
            ca.fb = fb;
            ca.tile_y0 = r->strip_y1 > r->strip_y0 ? r->strip_y0 : fb.min_y;
            ca.tile_y1 = r->strip_y1 > r->strip_y0 ? r->strip_y1 : fb.max_y;
            fill_destinations(r, current_target(r), ca);
            ca.clear_color = clear_color(r);
            ca.work_counter = r->counters.ptr + 12;
            wait_for_gather(r, r->stream); // (an empty frame is gathered as a whole frame: nothing was exported)
            r->stats.drawcall_count += (uint64_t)launch_composite(ca, r->stream);
        }
        if (r->borrowed_copies_pending && !r->pending.active) {
            // payload_persists uploads of a frame that ended without a verification wait
            PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
            r->borrowed_copies_pending = false;
        }
        r->stats.gpu_bytes_allocated = r->bytes_allocated;
        r->stats.gpu_bytes_committed = r->bytes_allocated;
    });
}

PFCudaStatus PFCudaRendererReadPixels(PFCudaRendererRef r, uint8_t *dst, size_t stride) {
    return guarded(r, [&]() {
        verify_pending(r);
        size_t row = (size_t)r->options.dest_size.x * 4;
        if (!dst || stride < row) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "bad destination / stride");
        wait_for_gather(r, r->stream); // the other ranks' strips
        PF_CUDA_CHECK(cudaMemcpy2DAsync(dst, stride, r->dest, r->dest_pitch, row, (size_t)r->options.dest_size.y,
                                        cudaMemcpyDeviceToHost, r->stream));
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    });
}

PFCudaStatus PFCudaRendererReadTexturePage(PFCudaRendererRef r, uint32_t page_id, uint8_t *dst, size_t stride,
                                           PFVector2I *size_out) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (page_id >= r->pages.size() || !r->pages[page_id]) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "texture page was never allocated");
        const PFCudaRenderer::TexturePage &page = *r->pages[page_id];
        if (size_out) *size_out = PFVector2I{page.width, page.height};
        if (!dst) return; // size query
        const size_t row = (size_t)page.width * 4;
        if (stride < row) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "stride too small");
        PF_CUDA_CHECK(cudaMemcpy2DAsync(dst, stride, page.pixels.ptr, row, row, (size_t)page.height, cudaMemcpyDeviceToHost, r->stream));
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    });
}

PFCudaStatus PFCudaRendererGetDestDevicePointer(PFCudaRendererRef r, uint64_t *device_ptr, size_t *pitch) {
    return guarded(r, [&]() {
        if (device_ptr) *device_ptr = (uint64_t)(uintptr_t)r->dest;
        if (pitch) *pitch = r->dest_pitch;
    });
}

PFCudaStatus PFCudaRendererSetDestDevicePointer(PFCudaRendererRef r, uint64_t device_ptr, size_t pitch) {
    return guarded(r, [&]() {
        verify_pending(r);
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        if (device_ptr == 0) {
            r->dest_external = false;
            allocate_dest(r);
        } else {
            if (pitch < (size_t)r->options.dest_size.x * 4) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "pitch too small");
            r->dest_external = true;
            r->dest = reinterpret_cast<uint8_t *>((uintptr_t)device_ptr);
            r->dest_pitch = pitch;
            r->dest_external_rows = r->options.dest_size.y; // the caller sized the buffer for the current frame
        }
    });
}

PFCudaStatus PFCudaIpcExport(uint64_t device_ptr, uint8_t handle_out[64], uint64_t *offset_out) {
    try {
        if (!device_ptr || !handle_out || !offset_out) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        PF_CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t)device_ptr)));
        memcpy(handle_out, &h, 64);
        // The handle names the whole allocation: report where the pointer sits inside it.
        // (driver entry point fetched at run time: the library must load on machines without libcuda)
        typedef CUresult (*GetAddressRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult query;
        PF_CUDA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &query));
        CUdeviceptr base = 0;
        size_t size = 0;
        if (!fn || reinterpret_cast<GetAddressRangeFn>(fn)(&base, &size, (CUdeviceptr)device_ptr) != CUDA_SUCCESS)
            throw Error(PF_CUDA_ERROR_CUDA, "cuMemGetAddressRange failed");
        *offset_out = (uint64_t)device_ptr - (uint64_t)base;
        return PF_CUDA_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.status;
    }
}

PFCudaStatus PFCudaRendererSetPeerDests(PFCudaRendererRef r, const uint8_t *handles, const uint64_t *offsets,
                                        int32_t count) {
    return guarded(r, [&]() {
        verify_pending(r);
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        close_peers(r);
        if (count < 0 || count > 7) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "at most 7 peers");
        for (int i = 0; i < count; i++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + (size_t)i * 64, 64);
            void *base = nullptr;
            PF_CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            r->peer_base[i] = base;
            r->peer_dest[i] = static_cast<uint8_t *>(base) + offsets[i];
            r->n_peers = i + 1;
        }
        if (count > 0 && !r->dest_external) r->dest_external_rows = r->options.dest_size.y;
    });
}

PFCudaStatus PFCudaRendererSetStream(PFCudaRendererRef r, uint64_t cuda_stream) {
    return guarded(r, [&]() {
        verify_pending(r);
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        r->stream = cuda_stream ? reinterpret_cast<cudaStream_t>((uintptr_t)cuda_stream) : r->own_stream;
    });
}

PFCudaStatus PFCudaRendererSynchronize(PFCudaRendererRef r) {
    return guarded(r, [&]() {
        verify_pending(r);
        wait_for_gather(r, r->stream);
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    });
}

void PFCudaStripOfRank(int32_t tile_rows, int32_t rank, int32_t world_size, int32_t *tile_y0, int32_t *tile_y1) {
    int32_t y0 = 0, y1 = 0;
    if (world_size > 0 && rank >= 0 && rank < world_size && tile_rows >= 0) strip_of_rank(tile_rows, rank, world_size, y0, y1);
    if (tile_y0) *tile_y0 = y0;
    if (tile_y1) *tile_y1 = y1;
}

PFCudaStatus PFCudaGatherCreateId(PFCudaGatherId *id_out) {
    try {
        if (!id_out) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null id");
        static_assert(sizeof(PFCudaGatherId) == 128, "ncclUniqueId size");
        const Nccl &n = nccl_or_throw();
        n.check(n.get_unique_id(id_out), "ncclGetUniqueId");
        return PF_CUDA_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return (PFCudaStatus)e.status;
    }
}

PFCudaStatus PFCudaRendererGatherInit(PFCudaRendererRef r, const PFCudaGatherId *id, int32_t rank, int32_t world_size) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (!id || world_size < 1 || rank < 0 || rank >= world_size)
            throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "bad gather id / rank / world size");
        if (r->dest_pitch != (size_t)r->options.dest_size.x * 4)
            throw Error(PF_CUDA_ERROR_UNSUPPORTED, "frame assembly needs a destination with contiguous rows");
        const Nccl &n = nccl_or_throw();
        gather_destroy(r);
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        void *comm = nullptr;
        n.check(n.comm_init_rank(&comm, world_size, *id, rank), "ncclCommInitRank");
        r->gather_comm = comm;
        r->gather_rank = rank;
        r->gather_world = world_size;
        PF_CUDA_CHECK(cudaStreamCreateWithFlags(&r->gather_stream, cudaStreamNonBlocking));
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&r->gather_ready, cudaEventDisableTiming));
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&r->gather_done, cudaEventDisableTiming));
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&r->gather_barrier_done, cudaEventDisableTiming));
        for (int q = 0; q < 2; q++) {
            PF_CUDA_CHECK(cudaEventCreateWithFlags(&r->push_done[q], cudaEventDisableTiming));
            r->push_recorded[q] = false;
        }
        const FbRect fb = framebuffer_tile_rect(r);
        strip_of_rank(fb.max_y - fb.min_y, rank, world_size, r->strip_y0, r->strip_y1);
        if (r->strip_y1 == r->strip_y0) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "more ranks than tile rows");

        // Tile mode: allocate the export region, exchange its IPC handle with the peers (an ncclAllGather of the 64-byte
        // handles), map theirs. Any failure (no peer access, more than 8 ranks, IPC refused) leaves FRAME mode on.
        r->gather_mode = PF_CUDA_GATHER_MODE_FRAME;
        r->gather_frame_serial = 0;
        PF_CUDA_CHECK(cudaMalloc((void **)&r->barrier_word, 64));
        PF_CUDA_CHECK(cudaMemset(r->barrier_word, 0, 64));
        if (world_size > 1 && world_size <= 8) {
            const int fb_w = fb.max_x - fb.min_x;
            int32_t max_rows = 0;
            for (int32_t g = 0; g < world_size; g++) {
                int32_t y0, y1;
                strip_of_rank(fb.max_y - fb.min_y, g, world_size, y0, y1);
                max_rows = std::max(max_rows, y1 - y0);
            }
            const size_t tiles = (size_t)fb_w * (size_t)max_rows, segments = (size_t)((fb_w + 31) / 32) * (size_t)max_rows;
            auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
            r->export_queue_off = 256;
            r->export_color_off = align(r->export_queue_off + tiles * 4);
            r->export_mask_off = align(r->export_color_off + segments * 32 * 4);
            r->export_blocks_off = align(r->export_mask_off + segments * 4);
            r->export_buffer_bytes = align(r->export_blocks_off + tiles * 1024);
            r->export_region.bytes_allocated = &r->bytes_allocated;
            r->export_region.ensure((size_t)world_size * 2 * r->export_buffer_bytes);
            PF_CUDA_CHECK(cudaMemset(r->export_region.ptr, 0, (size_t)world_size * 2 * r->export_buffer_bytes));
            cudaIpcMemHandle_t mine;
            uint8_t *handles_dev = nullptr;
            std::vector<cudaIpcMemHandle_t> handles(world_size);
            bool ok = cudaIpcGetMemHandle(&mine, r->export_region.ptr) == cudaSuccess;
            if (!ok) cudaGetLastError();
            // (collective either way: a rank whose export failed sends a zeroed handle and every rank falls back)
            if (!ok) memset(&mine, 0, sizeof(mine));
            PF_CUDA_CHECK(cudaMalloc((void **)&handles_dev, (size_t)world_size * 64));
            PF_CUDA_CHECK(cudaMemcpy(handles_dev + (size_t)rank * 64, &mine, 64, cudaMemcpyHostToDevice));
            n.check(n.all_gather(handles_dev + (size_t)rank * 64, handles_dev, 64, Nccl::UINT8, r->gather_comm, r->gather_stream),
                    "ncclAllGather (IPC handles)");
            PF_CUDA_CHECK(cudaStreamSynchronize(r->gather_stream));
            PF_CUDA_CHECK(cudaMemcpy(handles.data(), handles_dev, (size_t)world_size * 64, cudaMemcpyDeviceToHost));
            cudaFree(handles_dev);
            const cudaIpcMemHandle_t zero{};
            for (int32_t g = 0; g < world_size && ok; g++) {
                if (memcmp(&handles[g], &zero, sizeof(zero)) == 0) ok = false;
            }
            for (int32_t g = 0; g < world_size && ok; g++) {
                if (g == rank) {
                    r->peer_export[g] = r->export_region.ptr;
                    continue;
                }
                void *base = nullptr;
                if (cudaIpcOpenMemHandle(&base, handles[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = false;
                    break;
                }
                r->peer_export_base[g] = base;
                r->peer_export[g] = static_cast<uint8_t *>(base);
            }
            // every rank must end up in the same mode: agree on the minimum
            uint32_t vote = ok ? 1u : 0u, *vote_dev = r->barrier_word + 8;
            PF_CUDA_CHECK(cudaMemcpy(vote_dev, &vote, 4, cudaMemcpyHostToDevice));
            n.check(n.all_reduce(vote_dev, vote_dev, 1, Nccl::UINT32, Nccl::SUM, r->gather_comm, r->gather_stream), "ncclAllReduce (mode vote)");
            PF_CUDA_CHECK(cudaStreamSynchronize(r->gather_stream));
            PF_CUDA_CHECK(cudaMemcpy(&vote, vote_dev, 4, cudaMemcpyDeviceToHost));
            if (vote == (uint32_t)world_size) {
                r->gather_mode = PF_CUDA_GATHER_MODE_TILES;
            } else {
                for (int g = 0; g < 8; g++) {
                    if (r->peer_export_base[g]) cudaIpcCloseMemHandle(r->peer_export_base[g]);
                    r->peer_export_base[g] = nullptr;
                    r->peer_export[g] = nullptr;
                }
                r->export_region.release();
            }
        }
    });
}

PFCudaStatus PFCudaRendererGatherSetMode(PFCudaRendererRef r, int32_t mode) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (!r->gather_comm) throw Error(PF_CUDA_ERROR_PROTOCOL, "PFCudaRendererGatherSetMode before PFCudaRendererGatherInit");
        if (mode != PF_CUDA_GATHER_MODE_FRAME && mode != PF_CUDA_GATHER_MODE_TILES)
            throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "unknown gather mode");
        if (mode == PF_CUDA_GATHER_MODE_TILES && !r->peer_export[r->gather_rank])
            throw Error(PF_CUDA_ERROR_UNSUPPORTED, "tile mode needs the ranks' export regions mapped over CUDA IPC (not available here)");
        wait_for_gather(r, r->stream);
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        r->gather_mode = mode;
    });
}

PFCudaStatus PFCudaRendererGatherFrame(PFCudaRendererRef r) {
    return guarded(r, [&]() {
        if (!r->gather_comm) throw Error(PF_CUDA_ERROR_PROTOCOL, "PFCudaRendererGatherFrame before PFCudaRendererGatherInit");
        if (r->in_scene) throw Error(PF_CUDA_ERROR_PROTOCOL, "PFCudaRendererGatherFrame inside begin_scene / end_scene");
        const Nccl &n = Nccl::get();
        const FbRect fb = framebuffer_tile_rect(r);
        const int32_t rows = fb.max_y - fb.min_y, height = r->options.dest_size.y;
        const size_t pitch = r->dest_pitch;
        auto span = [&](int32_t rank, size_t &offset, size_t &bytes) { // the rank's strip in the frame, in bytes
            int32_t y0, y1;
            strip_of_rank(rows, rank, r->gather_world, y0, y1);
            const int32_t p0 = std::min(y0 * PF_TILE_HEIGHT, height), p1 = std::min(y1 * PF_TILE_HEIGHT, height);
            offset = (size_t)p0 * pitch, bytes = (size_t)(p1 - p0) * pitch;
        };
        // Ordered after this frame's compositing (and a gather still in flight on the same stream).
        PF_CUDA_CHECK(cudaEventRecord(r->gather_ready, r->stream));
        PF_CUDA_CHECK(cudaStreamWaitEvent(r->gather_stream, r->gather_ready, 0));
        if (r->gather_mode == PF_CUDA_GATHER_MODE_TILES && r->frame_exported) {
            // Push this strip's export into every peer's slot for this rank, barrier (every rank's push has landed), then
            // expand what the peers pushed into this rank's receive slots.
            {
                const size_t slot = ((size_t)r->gather_rank * 2 + (r->gather_frame_serial & 1)) * r->export_buffer_bytes;
                PushArgs push{};
                push.local = r->export_region.ptr + slot;
                for (int32_t g = 0; g < r->gather_world; g++)
                    if (g != r->gather_rank) push.remote[push.n_remote++] = r->peer_export[g] + slot;
                push.queue_off = r->export_queue_off, push.color_off = r->export_color_off;
                push.mask_off = r->export_mask_off, push.blocks_off = r->export_blocks_off;
                push.segments = (uint32_t)((fb.max_x - fb.min_x + 31) / 32) * (uint32_t)(r->strip_y1 - r->strip_y0);
                r->stats.drawcall_count += (uint64_t)launch_push_export(push, r->gather_stream);
                const int parity = (int)(r->gather_frame_serial & 1);
                PF_CUDA_CHECK(cudaEventRecord(r->push_done[parity], r->gather_stream));
                r->push_recorded[parity] = true;
            }
            n.check(n.all_reduce(r->barrier_word, r->barrier_word, 1, Nccl::UINT32, Nccl::SUM, r->gather_comm, r->gather_stream),
                    "ncclAllReduce (barrier)");
            PF_CUDA_CHECK(cudaEventRecord(r->gather_barrier_done, r->gather_stream));
            PullArgs pa{};
            for (int32_t g = 0; g < r->gather_world; g++) {
                if (g == r->gather_rank) continue;
                const uint8_t *buffer = r->export_region.ptr + ((size_t)g * 2 + (r->gather_frame_serial & 1)) * r->export_buffer_bytes;
                PullPeer &peer = pa.peers[pa.n_peers++];
                peer.alpha_count = reinterpret_cast<const uint32_t *>(buffer);
                peer.queue = reinterpret_cast<const uint32_t *>(buffer + r->export_queue_off);
                peer.solid_color = reinterpret_cast<const uint32_t *>(buffer + r->export_color_off);
                peer.solid_mask = reinterpret_cast<const uint32_t *>(buffer + r->export_mask_off);
                peer.blocks = buffer + r->export_blocks_off;
                strip_of_rank(rows, g, r->gather_world, peer.tile_y0, peer.tile_y1);
            }
            pa.fb = fb;
            pa.dest = r->dest;
            pa.dest_pitch = pitch;
            pa.dest_w = r->options.dest_size.x;
            pa.dest_h = height;
            r->stats.drawcall_count += (uint64_t)launch_pull_tiles(pa, r->gather_stream);
            PF_CUDA_CHECK(cudaEventRecord(r->gather_done, r->gather_stream));
            r->gather_in_flight = true;
            r->gather_last_was_tiles = true;
            r->gather_frame_serial++;
            return;
        }
        r->gather_last_was_tiles = false;
        size_t my_offset, my_bytes, offset0, bytes0;
        span(r->gather_rank, my_offset, my_bytes);
        span(0, offset0, bytes0);
        bool equal = true;
        for (int32_t g = 0; g < r->gather_world; g++) {
            size_t o, b;
            span(g, o, b);
            equal = equal && b == bytes0 && o == (size_t)g * bytes0;
        }
        if (equal) {
            n.check(n.all_gather(r->dest + my_offset, r->dest, my_bytes, Nccl::UINT8, r->gather_comm, r->gather_stream),
                    "ncclAllGather");
        } else {
            n.check(n.group_start(), "ncclGroupStart");
            for (int32_t g = 0; g < r->gather_world; g++) {
                size_t o, b;
                span(g, o, b);
                n.check(n.broadcast(r->dest + o, r->dest + o, b, Nccl::UINT8, g, r->gather_comm, r->gather_stream),
                        "ncclBroadcast");
            }
            n.check(n.group_end(), "ncclGroupEnd");
        }
        PF_CUDA_CHECK(cudaEventRecord(r->gather_done, r->gather_stream));
        r->gather_in_flight = true;
    });
}

PFCudaStatus PFCudaRendererGatherWait(PFCudaRendererRef r) {
    return guarded(r, [&]() { wait_for_gather(r, r->stream); });
}

PFCudaStatus PFCudaRendererGatherDestroy(PFCudaRendererRef r) {
    return guarded(r, [&]() {
        verify_pending(r);
        gather_destroy(r);
    });
}

PFCudaStatus PFCudaRendererSetStrip(PFCudaRendererRef r, int32_t tile_y0, int32_t tile_y1) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (tile_y1 < tile_y0) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "strip rows out of order");
        r->strip_y0 = tile_y0;
        r->strip_y1 = tile_y1;
    });
}

PFCudaStatus PFCudaRendererSetViewBox(PFCudaRendererRef r, const PFRectF *view_box) {
    return guarded(r, [&]() {
        if (!view_box) {
            r->has_view_box = false;
            return;
        }
        r->has_view_box = true;
        r->view_box = ViewBox{view_box->origin.x, view_box->origin.y, view_box->lower_right.x, view_box->lower_right.y};
    });
}

static PFCudaStatus forward_command(const PFRenderCommand *command, void *userdata) {
    return PFCudaRendererRenderCommand(static_cast<PFCudaRendererRef>(userdata), command);
}

// Scene::build_and_render (renderer/src/scene.rs:369-378).
PFCudaStatus PFSceneBuildAndRenderCuda(PFSceneRef scene, PFCudaRendererRef r, PFBuildOptionsRef options) {
    if (!scene || !r || !options) {
        set_last_error("PFSceneBuildAndRenderCuda: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    PFRectF vb;
    PFSceneGetViewBox(scene, &vb);
    PFCudaStatus st = PFCudaRendererSetViewBox(r, &vb);
    if (st != PF_CUDA_OK) return st;
    st = PFCudaRendererBeginScene(r);
    if (st != PF_CUDA_OK) return st;
    // The scene's segment arrays outlive the frame, so their upload does not have to be waited for
    // here: the scene itself waits (scene_note_borrowed) before it rewrites them.
    r->uploaded_scene_this_frame = false;
    pf::g_scene_payload_persists = true;
    pf::g_scene_strip[0] = r->strip_y0, pf::g_scene_strip[1] = r->strip_y1; // the builder skips paths outside the strip
    st = PFSceneBuild(scene, options, &r->sink_state, forward_command, r);
    pf::g_scene_strip[0] = pf::g_scene_strip[1] = 0;
    pf::g_scene_payload_persists = false;
    if (r->uploaded_scene_this_frame) {
        try {
            pf::scene_note_borrowed(scene, r->stream, r->ordinal);
        } catch (const Error &e) {
            set_last_error(e.what());
            if (st == PF_CUDA_OK) st = (PFCudaStatus)e.status;
        }
    }
    PFCudaStatus end = PFCudaRendererEndScene(r);
    return st != PF_CUDA_OK ? st : end;
}

// ---------------------------------------------------------------------------------------------
// SceneProxy (renderer/src/concurrent/scene_proxy.rs:35-157): the scene lives on a worker thread; replace_scene,
// set_view_box and build are messages the worker applies in order, render pumps the commands of the oldest queued
// build into the renderer. The reference moves owned commands through a channel; here a command's payload points into
// the scene's own arrays, so commands are handed over one at a time: the worker waits inside its listener until the
// caller has consumed the command. What runs beside the caller is the build up to each command.
// ---------------------------------------------------------------------------------------------
struct PFSceneProxy {
    enum Kind { REPLACE, SET_VIEW_BOX, BUILD, COPY };
    struct Msg {
        Kind kind;
        PFScene *scene;
        PFRectF view_box;
        PFBuildOptionsRef options;
        bool persists;
        int32_t strip[2];
    };
    PFScene *scene = nullptr;            // the worker's, once the thread runs
    PFRectF view_box{};                  // caller-side copy of the scene's view box (what the next build will see)
    std::deque<PFRectF> build_view_boxes; // ... as it was when each queued build was requested
    PFSceneSinkState sink_state{0, 0, 0}; // one sink per proxy (scene_proxy.rs:70)
    std::thread worker;
    std::mutex m;
    std::condition_variable cv;
    std::deque<Msg> inbox;
    bool quit = false;
    const PFRenderCommand *pending = nullptr; // hand-over of one command: set by the worker, consumed by the caller
    bool pending_ready = false, pending_done = false;
    PFCudaStatus pending_result = PF_CUDA_OK;
    bool build_done = false;             // the current build has ended; the worker goes on once the caller has seen it
    PFCudaStatus build_status = PF_CUDA_OK;
    std::string build_error;
    uint32_t builds_queued = 0;
    PFScene *copy_result = nullptr;
    bool copy_done = false;
};

namespace {

PFCudaStatus proxy_listener(const PFRenderCommand *command, void *userdata) {
    PFSceneProxy *p = static_cast<PFSceneProxy *>(userdata);
    std::unique_lock<std::mutex> lock(p->m);
    if (p->quit) return PF_CUDA_ERROR_PROTOCOL;
    p->pending = command;
    p->pending_ready = true, p->pending_done = false;
    p->cv.notify_all();
    p->cv.wait(lock, [&] { return p->pending_done || p->quit; });
    p->pending = nullptr;
    p->pending_ready = false;
    return p->pending_done ? p->pending_result : PF_CUDA_ERROR_PROTOCOL;
}

void proxy_thread(PFSceneProxy *p) {
    for (;;) {
        PFSceneProxy::Msg msg;
        {
            std::unique_lock<std::mutex> lock(p->m);
            p->cv.wait(lock, [&] { return p->quit || (!p->inbox.empty() && !p->build_done); });
            if (p->quit) return;
            msg = p->inbox.front();
            p->inbox.pop_front();
        }
        switch (msg.kind) {
        case PFSceneProxy::REPLACE:
            PFSceneDestroy(p->scene);
            p->scene = msg.scene;
            break;
        case PFSceneProxy::SET_VIEW_BOX: PFSceneSetViewBox(p->scene, &msg.view_box); break;
        case PFSceneProxy::COPY: {
            PFScene *copy = PFSceneClone(p->scene);
            std::lock_guard<std::mutex> lock(p->m);
            p->copy_result = copy, p->copy_done = true;
            p->cv.notify_all();
            break;
        }
        case PFSceneProxy::BUILD: {
            pf::g_scene_payload_persists = msg.persists;
            pf::g_scene_strip[0] = msg.strip[0], pf::g_scene_strip[1] = msg.strip[1];
            const PFCudaStatus st = PFSceneBuild(p->scene, msg.options, &p->sink_state, proxy_listener, p);
            pf::g_scene_strip[0] = pf::g_scene_strip[1] = 0;
            pf::g_scene_payload_persists = false;
            const std::string error = st != PF_CUDA_OK ? std::string(PFCudaGetLastError()) : std::string();
            PFBuildOptionsDestroy(msg.options);
            std::lock_guard<std::mutex> lock(p->m);
            p->build_done = true, p->build_status = st, p->build_error = error;
            p->cv.notify_all();
            break;
        }
        }
    }
}

bool proxy_send(PFSceneProxy *p, const PFSceneProxy::Msg &msg) {
    std::lock_guard<std::mutex> lock(p->m);
    if (p->quit) return false;
    p->inbox.push_back(msg);
    p->cv.notify_all();
    return true;
}

// The commands of the oldest queued build, one at a time, to `listener`. Returns once the build has ended; with
// acknowledge = false the worker stays parked (the caller still has something to do with the idle scene).
PFCudaStatus proxy_receive(PFSceneProxy *p, PFRenderCommandListenerFn listener, void *userdata, bool acknowledge) {
    std::unique_lock<std::mutex> lock(p->m);
    if (p->builds_queued == 0) {
        set_last_error("scene proxy: render without a build");
        return PF_CUDA_ERROR_PROTOCOL;
    }
    PFCudaStatus aborted = PF_CUDA_OK;
    std::string abort_error;
    for (;;) {
        p->cv.wait(lock, [&] { return p->pending_ready || p->build_done; });
        if (!p->pending_ready) break; // the build has ended
        const PFRenderCommand *command = p->pending;
        p->pending_ready = false;
        lock.unlock();
        const PFCudaStatus st = listener(command, userdata);
        if (st != PF_CUDA_OK && aborted == PF_CUDA_OK) aborted = st, abort_error = PFCudaGetLastError();
        lock.lock();
        p->pending_result = st, p->pending_done = true;
        p->cv.notify_all();
    }
    const PFCudaStatus status = aborted != PF_CUDA_OK ? aborted : p->build_status;
    if (status != PF_CUDA_OK) set_last_error(aborted != PF_CUDA_OK ? abort_error : p->build_error);
    if (acknowledge) {
        p->build_done = false;
        p->builds_queued--;
        p->cv.notify_all();
    }
    return status;
}

void proxy_acknowledge(PFSceneProxy *p) {
    std::lock_guard<std::mutex> lock(p->m);
    p->build_done = false;
    p->builds_queued--;
    p->cv.notify_all();
}

PFCudaStatus proxy_queue_build(PFSceneProxy *p, PFBuildOptionsRef options, bool persists, int32_t y0, int32_t y1) {
    if (!p || !options) {
        set_last_error("scene proxy: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    PFSceneProxy::Msg msg{PFSceneProxy::BUILD, nullptr, PFRectF{}, PFBuildOptionsClone(options), persists, {y0, y1}};
    {
        std::lock_guard<std::mutex> lock(p->m);
        p->builds_queued++;
        p->build_view_boxes.push_back(p->view_box);
    }
    if (!proxy_send(p, msg)) {
        PFBuildOptionsDestroy(msg.options);
        return PF_CUDA_ERROR_PROTOCOL;
    }
    return PF_CUDA_OK;
}

} // namespace

PFSceneProxyRef PFSceneProxyCreateFromScene(PFSceneRef scene) {
    if (!scene) {
        set_last_error("PFSceneProxyCreateFromScene: null scene");
        return nullptr;
    }
    PFSceneProxy *p = new PFSceneProxy();
    p->scene = scene;
    PFSceneGetViewBox(scene, &p->view_box);
    p->worker = std::thread(proxy_thread, p);
    return p;
}

void PFSceneProxyDestroy(PFSceneProxyRef p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(p->m);
        p->quit = true;
        p->cv.notify_all();
    }
    p->worker.join(); // (a build in flight is aborted: its listener returns an error)
    for (PFSceneProxy::Msg &msg : p->inbox) {
        if (msg.kind == PFSceneProxy::REPLACE) PFSceneDestroy(msg.scene);
        if (msg.kind == PFSceneProxy::BUILD) PFBuildOptionsDestroy(msg.options);
    }
    PFSceneDestroy(p->scene);
    delete p;
}

PFCudaStatus PFSceneProxyReplaceScene(PFSceneProxyRef p, PFSceneRef new_scene) {
    if (!p || !new_scene) {
        set_last_error("PFSceneProxyReplaceScene: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    PFRectF vb;
    PFSceneGetViewBox(new_scene, &vb);
    {
        std::lock_guard<std::mutex> lock(p->m);
        p->view_box = vb;
    }
    return proxy_send(p, PFSceneProxy::Msg{PFSceneProxy::REPLACE, new_scene, vb, nullptr, false, {0, 0}}) ? PF_CUDA_OK : PF_CUDA_ERROR_PROTOCOL;
}

PFCudaStatus PFSceneProxySetViewBox(PFSceneProxyRef p, const PFRectF *view_box) {
    if (!p || !view_box) {
        set_last_error("PFSceneProxySetViewBox: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    {
        std::lock_guard<std::mutex> lock(p->m);
        p->view_box = *view_box;
    }
    return proxy_send(p, PFSceneProxy::Msg{PFSceneProxy::SET_VIEW_BOX, nullptr, *view_box, nullptr, false, {0, 0}}) ? PF_CUDA_OK : PF_CUDA_ERROR_PROTOCOL;
}

PFCudaStatus PFSceneProxyBuild(PFSceneProxyRef p, PFBuildOptionsRef options) {
    return proxy_queue_build(p, options, false, 0, 0);
}

PFCudaStatus PFSceneProxyReceive(PFSceneProxyRef p, PFRenderCommandListenerFn listener, void *userdata) {
    if (!p || !listener) {
        set_last_error("PFSceneProxyReceive: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    {
        std::lock_guard<std::mutex> lock(p->m);
        if (!p->build_view_boxes.empty()) p->build_view_boxes.pop_front();
    }
    return proxy_receive(p, listener, userdata, true);
}

PFCudaStatus PFSceneProxyRenderCuda(PFSceneProxyRef p, PFCudaRendererRef r) {
    if (!p || !r) {
        set_last_error("PFSceneProxyRenderCuda: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    PFRectF vb;
    {
        std::lock_guard<std::mutex> lock(p->m);
        if (p->builds_queued == 0 || p->build_view_boxes.empty()) {
            set_last_error("scene proxy: render without a build");
            return PF_CUDA_ERROR_PROTOCOL;
        }
        vb = p->build_view_boxes.front();
        p->build_view_boxes.pop_front();
    }
    PFCudaStatus st = PFCudaRendererSetViewBox(r, &vb);
    if (st == PF_CUDA_OK) st = PFCudaRendererBeginScene(r);
    if (st != PF_CUDA_OK) { // the build still has to be consumed: drop its commands
        proxy_receive(p, [](const PFRenderCommand *, void *) -> PFCudaStatus { return PF_CUDA_ERROR_PROTOCOL; }, nullptr, true);
        return st;
    }
    r->uploaded_scene_this_frame = false;
    st = proxy_receive(p, forward_command, r, false);
    if (r->uploaded_scene_this_frame) { // (the worker is parked until the acknowledgement: the scene is idle)
        try {
            pf::scene_note_borrowed(p->scene, r->stream, r->ordinal);
        } catch (const Error &e) {
            set_last_error(e.what());
            if (st == PF_CUDA_OK) st = (PFCudaStatus)e.status;
        }
    }
    proxy_acknowledge(p);
    const PFCudaStatus end = PFCudaRendererEndScene(r);
    return st != PF_CUDA_OK ? st : end;
}

PFCudaStatus PFSceneProxyBuildAndRenderCuda(PFSceneProxyRef p, PFCudaRendererRef r, PFBuildOptionsRef options) {
    if (!p || !r || !options) {
        set_last_error("PFSceneProxyBuildAndRenderCuda: null argument");
        return PF_CUDA_ERROR_INVALID_ARGUMENT;
    }
    const PFCudaStatus st = proxy_queue_build(p, options, true, r->strip_y0, r->strip_y1);
    return st != PF_CUDA_OK ? st : PFSceneProxyRenderCuda(p, r);
}

PFSceneRef PFSceneProxyCopyScene(PFSceneProxyRef p) {
    if (!p) return nullptr;
    {
        std::lock_guard<std::mutex> lock(p->m);
        if (p->builds_queued != 0) { // (the worker hands its commands over one at a time: it would never get to the copy)
            set_last_error("scene proxy: copy_scene with a build still to be rendered");
            return nullptr;
        }
        p->copy_done = false;
    }
    if (!proxy_send(p, PFSceneProxy::Msg{PFSceneProxy::COPY, nullptr, PFRectF{}, nullptr, false, {0, 0}})) return nullptr;
    std::unique_lock<std::mutex> lock(p->m);
    p->cv.wait(lock, [&] { return p->copy_done; });
    PFScene *copy = p->copy_result;
    p->copy_result = nullptr;
    return copy;
}

PFCudaStatus PFCudaRendererGetStats(PFCudaRendererRef r, PFCudaRenderStats *stats) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (!stats) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null stats");
        if (r->stats.fill_count == 0 && r->batches_drawn == 1 && !r->in_scene && r->last_batch.n_tiles) {
            // RenderStats.fill_count (all fills, before occlusion culling) is not needed to render:
            // summed on demand from the per-tile counts.
            unsigned long long *total = reinterpret_cast<unsigned long long *>(r->counters.ptr + 8);
            PF_CUDA_CHECK(cudaMemsetAsync(total, 0, sizeof(*total), r->stream));
            launch_sum_fill_counts(r->tile_word.ptr, r->last_batch.n_tiles, total, r->stream);
            unsigned long long host_total = 0;
            PF_CUDA_CHECK(cudaMemcpyAsync(&host_total, total, sizeof(host_total), cudaMemcpyDeviceToHost, r->stream));
            PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
            r->stats.fill_count = host_total;
        }
        *stats = r->stats;
        if (r->debug_lists && r->batches_drawn > 0 && !r->in_scene) {
            ensure_alpha_ids(r);
            stats->alpha_tile_count = r->last_alpha_tiles + ((r->cache.has_clips && r->clip.valid) ? r->clip.n_alpha : 0);
        }
    });
}

PFCudaStatus PFCudaRendererSetDeferredVerification(PFCudaRendererRef r, int32_t enabled) {
    return guarded(r, [&]() {
        verify_pending(r);
        r->deferred_verify = enabled != 0;
    });
}

PFCudaStatus PFCudaRendererSetTimingEnabled(PFCudaRendererRef r, int32_t enabled) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (enabled && !r->timing) r->times_total = PFCudaRenderTime{}, r->times_batches = 0;
        r->timing = enabled != 0;
    });
}

PFCudaStatus PFCudaRendererGetAccumulatedTimes(PFCudaRendererRef r, PFCudaRenderTime *times, uint32_t *batches) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (!times || !batches) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null argument");
        *times = r->times_total;
        *batches = r->times_batches;
    });
}

PFCudaStatus PFCudaRendererGetTimes(PFCudaRendererRef r, PFCudaRenderTime *times) {
    return guarded(r, [&]() {
        verify_pending(r);
        if (!times) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "null times");
        *times = r->times;
    });
}

PFCudaStatus PFCudaRendererSetDebugListsEnabled(PFCudaRendererRef r, int32_t enabled) {
    return guarded(r, [&]() {
        verify_pending(r);
        r->debug_lists = enabled != 0;
        r->cache.counts_valid = false; // the dump buffers must be sized on the next frame
    });
}

int64_t PFCudaRendererDebugCopyLines(PFCudaRendererRef r, float *out_lines, uint32_t *out_paths, size_t cap) {
    return guarded_count(r, [&]() -> int64_t {
        verify_pending(r);
        size_t n = r->last_lines, m = n < cap ? n : cap;
        if (out_lines && m) PF_CUDA_CHECK(cudaMemcpyAsync(out_lines, r->lines.ptr, m * sizeof(float4), cudaMemcpyDeviceToHost, r->stream));
        if (out_paths && m) PF_CUDA_CHECK(cudaMemcpyAsync(out_paths, r->line_path.ptr, m * 4, cudaMemcpyDeviceToHost, r->stream));
        PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        return (int64_t)n;
    });
}

int64_t PFCudaRendererDebugCopyFills(PFCudaRendererRef r, PFFill *out, size_t cap) {
    return guarded_count(r, [&]() -> int64_t {
        verify_pending(r);
        // With clipped paths the clip paths' fills come first (SequentialExecutor order: clip paths, then draw paths).
        const size_t n_clip = (r->cache.has_clips && r->clip.valid) ? r->clip.n_fills : 0;
        size_t n = n_clip + r->last_fills, m = n < cap ? n : cap;
        if (out && m) {
            ensure_alpha_ids(r);
            r->dump_out.ensure(n * sizeof(PFFill) + 16, 1.25);
            if (n_clip)
                PF_CUDA_CHECK(cudaMemcpyAsync(r->dump_out.ptr, r->clip.fill_records.ptr, n_clip * sizeof(PFFill),
                                              cudaMemcpyDeviceToDevice, r->stream));
            launch_dump_fills(r->last_fills, r->fills_emit.ptr, r->tile_alpha_id.ptr,
                              r->dump_out.ptr + n_clip * sizeof(PFFill), r->stream);
            PF_CUDA_CHECK(cudaMemcpyAsync(out, r->dump_out.ptr, m * sizeof(PFFill), cudaMemcpyDeviceToHost, r->stream));
            PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        }
        return (int64_t)n;
    });
}

int64_t PFCudaRendererDebugCopyTiles(PFCudaRendererRef r, PFTileObjectPrimitive *out, size_t cap) {
    return guarded_count(r, [&]() -> int64_t {
        verify_pending(r);
        ensure_alpha_ids(r);
        const BatchDev &b = r->last_batch;
        const bool clipped = r->cache.has_clips && r->clip.valid;
        // tile_fb is free again after the sort stage: reuse it for the non-empty flags.
        r->tile_pos.ensure(b.n_tiles + 1, 1.25);
        launch_dump_tile_flags(b.n_tiles, r->tile_word.ptr, r->tile_fb.ptr, r->stream);
        exclusive_scan(LoadU32{r->tile_fb.ptr}, r->tile_pos.ptr, b.n_tiles, r->counters.ptr + 4, r->scan_scratch, r->stream);
        size_t n = b.n_tiles ? read_counter(r, 4) : 0;
        size_t m = n < cap ? n : cap;
        if (out && m) {
            r->dump_out.ensure(n * sizeof(PFTileObjectPrimitive) + 16, 1.25);
            launch_dump_tiles(b, r->tile_word.ptr, r->tile_alpha_id.ptr, r->tile_fb.ptr, r->tile_pos.ptr, r->dump_out.ptr,
                              clipped ? r->tile_clip.ptr : nullptr, clipped ? r->clip.tile_word.ptr : nullptr,
                              clipped ? r->clip.tile_alpha_id.ptr : nullptr, r->stream);
            PF_CUDA_CHECK(cudaMemcpyAsync(out, r->dump_out.ptr, m * sizeof(PFTileObjectPrimitive), cudaMemcpyDeviceToHost, r->stream));
            PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        }
        return (int64_t)n;
    });
}

// The Clip records of the D3D9 batch (gpu_data.rs Clip, builder.rs:1031-1040), in tile order.
int64_t PFCudaRendererDebugCopyClips(PFCudaRendererRef r, PFClip *out, size_t cap) {
    return guarded_count(r, [&]() -> int64_t {
        verify_pending(r);
        if (!(r->cache.has_clips && r->clip.valid)) return 0;
        ensure_alpha_ids(r);
        const BatchDev &b = r->last_batch;
        r->tile_pos.ensure(b.n_tiles + 1, 1.25);
        launch_dump_clip_flags(b.n_tiles, r->tile_clip.ptr, r->tile_fb.ptr, r->stream);
        exclusive_scan(LoadU32{r->tile_fb.ptr}, r->tile_pos.ptr, b.n_tiles, r->counters.ptr + 4, r->scan_scratch, r->stream);
        size_t n = b.n_tiles ? read_counter(r, 4) : 0;
        size_t m = n < cap ? n : cap;
        if (out && m) {
            r->dump_out.ensure(n * sizeof(PFClip) + 16, 1.25);
            launch_dump_clips(b.n_tiles, r->tile_word.ptr, r->tile_alpha_id.ptr, r->tile_clip.ptr, r->clip.tile_word.ptr,
                              r->clip.tile_alpha_id.ptr, r->tile_fb.ptr, r->tile_pos.ptr, r->dump_out.ptr, r->stream);
            PF_CUDA_CHECK(cudaMemcpyAsync(out, r->dump_out.ptr, m * sizeof(PFClip), cudaMemcpyDeviceToHost, r->stream));
            PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        }
        return (int64_t)n;
    });
}

int64_t PFCudaRendererDebugCopyZBuffer(PFCudaRendererRef r, int32_t *out, size_t cap, int32_t rect_out[4]) {
    return guarded_count(r, [&]() -> int64_t {
        verify_pending(r);
        const FbRect &fb = r->last_fb;
        if (rect_out) {
            rect_out[0] = fb.min_x, rect_out[1] = fb.min_y, rect_out[2] = fb.max_x, rect_out[3] = fb.max_y;
        }
        size_t n = (size_t)(fb.max_x - fb.min_x) * (size_t)(fb.max_y - fb.min_y), m = n < cap ? n : cap;
        if (out && m) {
            PF_CUDA_CHECK(cudaMemcpyAsync(out, r->z_buffer.ptr, m * 4, cudaMemcpyDeviceToHost, r->stream));
            PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        }
        return (int64_t)n;
    });
}

int64_t PFCudaRendererDebugCopyAlphaMasks(PFCudaRendererRef r, float *out, size_t cap_tiles) {
    return guarded_count(r, [&]() -> int64_t {
        verify_pending(r);
        if (r->cache.has_clips && r->clip.valid)
            throw Error(PF_CUDA_ERROR_UNSUPPORTED, "alpha mask dumps do not cover clipped paths");
        ensure_alpha_ids(r);
        size_t n = r->last_alpha_tiles;
        if (out && n) {
            if (cap_tiles < n) throw Error(PF_CUDA_ERROR_INVALID_ARGUMENT, "alpha mask buffer too small");
            r->dump_out.ensure(n * 256 * sizeof(float) + 16, 1.25);
            launch_alpha_masks(r->last_batch.n_tiles, r->tile_word.ptr, r->tile_fill_pos.ptr, r->tile_alpha_id.ptr,
                               r->fills.ptr, r->lut_tex, reinterpret_cast<float *>(r->dump_out.ptr), r->stream);
            PF_CUDA_CHECK(cudaMemcpyAsync(out, r->dump_out.ptr, n * 256 * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
            PF_CUDA_CHECK(cudaStreamSynchronize(r->stream));
        }
        return (int64_t)n;
    });
}

} // extern "C"
