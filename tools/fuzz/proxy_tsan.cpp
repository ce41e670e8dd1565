// tools/fuzz/proxy_tsan.cpp — the scene proxy (worker thread, one-command hand-over, teardown with a build in flight) under
// ThreadSanitizer, without a GPU: commands are consumed by a listener (PFSceneProxyReceive). See run.sh.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "../../include/pf_cuda.h"

static PFCudaStatus count_tiles(const PFRenderCommand *c, void *ud) {
    if (c->kind == PF_RENDER_COMMAND_DRAW_TILES_D3D11) *(uint32_t *)ud += c->u.draw_tiles_d3d11.tile_batch_data.tile_count;
    return PF_CUDA_OK;
}
static PFCudaStatus fail_on_draw(const PFRenderCommand *c, void *) {
    return c->kind == PF_RENDER_COMMAND_DRAW_TILES_D3D11 ? PF_CUDA_ERROR_UNSUPPORTED : PF_CUDA_OK;
}

static PFSceneRef make_scene(unsigned seed, int n_paths) {
    unsigned r = seed;
    auto rnd = [&]() { r = r * 1664525u + 1013904223u; return (r >> 8) & 0xffff; };
    PFSceneRef s = PFSceneCreate();
    PFRectF vb{{0, 0}, {1024, 1024}};
    PFSceneSetViewBox(s, &vb);
    PFColorU c{1, 2, 3, 255};
    uint16_t paint = PFScenePushPaint(s, &c);
    std::vector<PFVector2F> pts; std::vector<uint8_t> fl; std::vector<uint32_t> off{0}, poff{0};
    std::vector<uint16_t> paints(n_paths, paint); std::vector<uint8_t> rules(n_paths, 0);
    for (int p = 0; p < n_paths; p++) {
        float cx = rnd() % 1024, cy = rnd() % 1024;
        for (int i = 0; i < 4; i++) { pts.push_back(PFVector2F{cx + (float)(rnd() % 40), cy + (float)(rnd() % 40)}); fl.push_back(0); }
        off.push_back((uint32_t)pts.size());
        poff.push_back((uint32_t)off.size() - 1);
    }
    PFScenePushDrawPaths(s, pts.data(), fl.data(), pts.size(), off.data(), off.size() - 1, poff.data(), n_paths, paints.data(), rules.data(), nullptr);
    return s;
}

static void drive(unsigned seed, uint32_t *tiles) {
    PFSceneProxyRef p = PFSceneProxyCreateFromScene(make_scene(seed, 5000));
    PFBuildOptionsRef o = PFBuildOptionsCreate();
    for (int round = 0; round < 4; round++) {
        PFSceneProxyBuild(p, o);
        PFRectF vb{{0, 0}, {512.0f + 64.0f * round, 512}};
        PFSceneProxySetViewBox(p, &vb);            // queued behind the build
        PFSceneProxyBuild(p, o);
        { PFCudaStatus st_ = PFSceneProxyReceive(p, count_tiles, tiles); if (st_ != PF_CUDA_OK) { fprintf(stderr, "receive: %d %s (line %d)\n", (int)st_, PFCudaGetLastError(), __LINE__); abort(); } }
        { PFCudaStatus st_ = PFSceneProxyReceive(p, count_tiles, tiles); if (st_ != PF_CUDA_OK) { fprintf(stderr, "receive: %d %s (line %d)\n", (int)st_, PFCudaGetLastError(), __LINE__); abort(); } }
        PFSceneRef copy = PFSceneProxyCopyScene(p);
        if (!copy) { fprintf(stderr, "copy: %s\n", PFCudaGetLastError()); abort(); }
        PFSceneDestroy(copy);
        if (round == 1) PFSceneProxyReplaceScene(p, make_scene(seed + 100, 2000));
    }
    PFSceneProxyBuild(p, o);                        // a listener that refuses: the build is aborted, the proxy lives on
    if (PFSceneProxyReceive(p, fail_on_draw, nullptr) != PF_CUDA_ERROR_UNSUPPORTED) abort();
    PFSceneProxyBuild(p, o);
    { PFCudaStatus st_ = PFSceneProxyReceive(p, count_tiles, tiles); if (st_ != PF_CUDA_OK) { fprintf(stderr, "receive: %d %s (line %d)\n", (int)st_, PFCudaGetLastError(), __LINE__); abort(); } }
    PFSceneProxyBuild(p, o);                        // torn down with a build in flight
    PFSceneProxyBuild(p, o);
    PFBuildOptionsDestroy(o);
    PFSceneProxyDestroy(p);
}

int main() {
    uint32_t a = 0, b = 0;
    std::thread t(drive, 1u, &a);
    drive(2u, &b);
    t.join();
    printf("ok %u %u\n", a, b);
    return 0;
}
