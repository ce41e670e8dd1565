// tools/fuzz/scene_fuzz.cpp — random scenes (NaN / infinite / huge coordinates, stray control flags, empty contours, clip paths, transforms,
// dilations) through PFSceneBuild and PFSceneBuildForStrip under ASan + UBSan; the listener reads every payload byte and re-derives the tile count (see run.sh).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/pf_cuda.h"
namespace pf { void set_last_error(const std::string &) {} }
static uint64_t sink_sum = 0;
template <typename T> static void touch(const T *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n * sizeof(T); i++) sink_sum += b[i];
}
static PFCudaStatus listener(const PFRenderCommand *c, void *) {
    switch (c->kind) {
    case PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11: {
        const auto &u = c->u.upload_scene_d3d11;
        touch(u.draw_segments.points, u.draw_segments.point_count);
        touch(u.draw_segments.indices, u.draw_segments.index_count);
        touch(u.clip_segments.points, u.clip_segments.point_count);
        touch(u.clip_segments.indices, u.clip_segments.index_count);
        // every index must address the point array with room for its curve
        for (size_t i = 0; i < u.draw_segments.index_count; i++) {
            uint32_t first = u.draw_segments.indices[i].first_point_index, fl = u.draw_segments.indices[i].flags;
            uint32_t need = (fl & PF_CURVE_IS_CUBIC) ? 4 : (fl & PF_CURVE_IS_QUADRATIC) ? 3 : 2;
            if ((size_t)first + need > u.draw_segments.point_count) { printf("BAD index %zu\n", i); abort(); }
        }
        break;
    }
    case PF_RENDER_COMMAND_UPLOAD_TEXTURE_METADATA:
        touch(c->u.upload_texture_metadata.entries, c->u.upload_texture_metadata.entry_count);
        break;
    case PF_RENDER_COMMAND_PREPARE_CLIP_TILES_D3D11:
    case PF_RENDER_COMMAND_DRAW_TILES_D3D11: {
        const PFTileBatchDataD3D11 &b = c->kind == PF_RENDER_COMMAND_DRAW_TILES_D3D11 ? c->u.draw_tiles_d3d11.tile_batch_data
                                                                                       : c->u.prepare_clip_tiles_d3d11.batch;
        touch(b.prepare_info.propagate_metadata, b.path_count);
        touch(b.prepare_info.dice_metadata, b.path_count);
        touch(b.prepare_info.tile_path_info, b.path_count);
        uint64_t tiles = 0;
        for (uint32_t i = 0; i < b.path_count; i++) {
            const PFRectI &r = b.prepare_info.propagate_metadata[i].tile_rect;
            if (r.lower_right.x < r.origin.x || r.lower_right.y < r.origin.y) { printf("BAD rect\n"); abort(); }
            tiles += (uint64_t)(r.lower_right.x - r.origin.x) * (uint64_t)(r.lower_right.y - r.origin.y);
        }
        if (tiles != b.tile_count) { printf("BAD tile count %llu vs %u\n", (unsigned long long)tiles, b.tile_count); abort(); }
        break;
    }
    default: break;
    }
    return PF_CUDA_OK;
}
static float rnd_coord() {
    int k = rand() % 40;
    if (k == 0) return NAN;
    if (k == 1) return INFINITY;
    if (k == 2) return -INFINITY;
    if (k == 3) return 1e30f;
    if (k == 4) return -3e38f;
    if (k == 5) return 0.0f;
    return (float)(rand() % 4000 - 1000) / 7.0f;
}
int main() {
    srand(3);
    for (int it = 0; it < 3000; it++) {
        PFSceneRef s = PFSceneCreate();
        PFRectF vb{{0, 0}, {(float)(16 + rand() % 600), (float)(16 + rand() % 600)}};
        if (it % 50 == 0) vb.lower_right.x = 0;  // empty view box
        PFSceneSetViewBox(s, &vb);
        std::vector<uint16_t> paints;
        for (int p = 0; p < 1 + rand() % 4; p++) {
            PFColorU c{(uint8_t)rand(), (uint8_t)rand(), (uint8_t)rand(), (uint8_t)(rand() % 2 ? 255 : rand())};
            paints.push_back(PFScenePushPaint(s, &c));
        }
        int n_clips = rand() % 3, n_paths = rand() % 12;
        auto make = [&](std::vector<PFVector2F> &pts, std::vector<uint8_t> &fl, std::vector<uint32_t> &off) {
            off.push_back(0);
            int contours = rand() % 4;
            for (int c = 0; c < contours; c++) {
                int n = rand() % 9;
                for (int i = 0; i < n; i++) {
                    pts.push_back(PFVector2F{rnd_coord(), rnd_coord()});
                    int f = rand() % 6;
                    fl.push_back(f == 0 ? 1 : f == 1 ? 2 : f == 2 ? 3 : 0);  // stray control flags included
                }
                off.push_back((uint32_t)pts.size());
            }
        };
        for (int c = 0; c < n_clips; c++) {
            std::vector<PFVector2F> pts; std::vector<uint8_t> fl; std::vector<uint32_t> off;
            make(pts, fl, off);
            PFScenePushClipPath(s, pts.data(), fl.data(), off.data(), (uint32_t)off.size() - 1, rand() % 2, PF_CLIP_PATH_NONE);
        }
        for (int p = 0; p < n_paths; p++) {
            std::vector<PFVector2F> pts; std::vector<uint8_t> fl; std::vector<uint32_t> off;
            make(pts, fl, off);
            uint32_t clip = (n_clips && rand() % 3 == 0) ? (uint32_t)(rand() % n_clips) : PF_CLIP_PATH_NONE;
            PFScenePushDrawPath(s, pts.data(), fl.data(), off.data(), (uint32_t)off.size() - 1, paints[rand() % paints.size()],
                                rand() % 2, 0, clip);
        }
        PFSceneSinkState sink{0, 0, 0};
        for (int build = 0; build < 3; build++) {
            PFBuildOptionsRef o = PFBuildOptionsCreate();
            if (rand() % 2) {
                PFTransform2F t{{(float)(rand() % 300) / 100.0f, (float)(rand() % 100) / 100.0f, (float)(rand() % 100) / -100.0f,
                                 (float)(rand() % 300) / 100.0f}, {rnd_coord(), (float)(rand() % 50)}};
                PFBuildOptionsSetTransform(o, PFRenderTransformCreate2D(&t));
            }
            if (rand() % 2) {
                PFVector2F d{(float)(rand() % 30) / 10.0f, rand() % 9 == 0 ? NAN : (float)(rand() % 30) / 10.0f};
                PFBuildOptionsSetDilation(o, &d);
            }
            PFSceneBuild(s, o, &sink, listener, nullptr);
            {   // the same for a renderer that owns a strip of tile rows (multi-GPU): rows inside, across and below the frame
                const int32_t y0 = rand() % 40 - 2, y1 = y0 + rand() % 24;
                PFSceneSinkState strip_sink{0, 0, 0};
                PFSceneBuildForStrip(s, o, &strip_sink, listener, nullptr, y0, y1);
            }
            PFBuildOptionsDestroy(o);
        }
        PFSceneDestroy(s);
    }
    printf("ok %llu\n", (unsigned long long)sink_sum);
}
