// tools/fuzz/scene_tsan.cpp — three threads building 30k-path scenes at once (shared worker pool, two-level scans, prepare_paths) under
// ThreadSanitizer (see run.sh).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <thread>
#include <vector>
#include "../../include/pf_cuda.h"
namespace pf { void set_last_error(const std::string &) {} }
static PFCudaStatus listener(const PFRenderCommand *c, void *ud) {
    if (c->kind == PF_RENDER_COMMAND_DRAW_TILES_D3D11) *(uint32_t *)ud += c->u.draw_tiles_d3d11.tile_batch_data.tile_count;
    return PF_CUDA_OK;
}
static void worker(int seed, uint32_t *out) {
    unsigned r = seed;
    auto rnd = [&]() { r = r * 1664525u + 1013904223u; return (r >> 8) & 0xffff; };
    PFSceneRef s = PFSceneCreate();
    PFRectF vb{{0, 0}, {2048, 2048}};
    PFSceneSetViewBox(s, &vb);
    PFColorU c{1, 2, 3, 255};
    uint16_t paint = PFScenePushPaint(s, &c);
    const int n_paths = 30000;
    std::vector<PFVector2F> pts; std::vector<uint8_t> fl; std::vector<uint32_t> off{0}, poff{0};
    std::vector<uint16_t> paints(n_paths, paint); std::vector<uint8_t> rules(n_paths, 0);
    for (int p = 0; p < n_paths; p++) {
        float cx = rnd() % 2048, cy = rnd() % 2048;
        for (int i = 0; i < 4; i++) { pts.push_back(PFVector2F{cx + (float)(rnd() % 40), cy + (float)(rnd() % 40)}); fl.push_back(0); }
        off.push_back((uint32_t)pts.size());
        poff.push_back((uint32_t)off.size() - 1);
    }
    PFScenePushDrawPaths(s, pts.data(), fl.data(), pts.size(), off.data(), off.size() - 1, poff.data(), n_paths, paints.data(), rules.data(), nullptr);
    PFSceneSinkState sink{0, 0, 0};
    for (int b = 0; b < 6; b++) {
        PFBuildOptionsRef o = PFBuildOptionsCreate();
        PFVector2F d{b % 2 ? 0.5f : 0.0f, b % 2 ? 0.25f * b : 0.0f};
        PFBuildOptionsSetDilation(o, &d);
        PFTransform2F t{{1.0f + 0.1f * b, 0, 0, 1.0f}, {0, 0}};
        PFBuildOptionsSetTransform(o, PFRenderTransformCreate2D(&t));
        PFSceneBuild(s, o, &sink, listener, out);
        PFBuildOptionsDestroy(o);
    }
    PFSceneDestroy(s);
}
int main() {
    uint32_t a = 0, b = 0, c = 0;
    std::thread t1(worker, 1, &a), t2(worker, 2, &b);
    worker(3, &c);
    t1.join(); t2.join();
    printf("ok %u %u %u\n", a, b, c);
}
