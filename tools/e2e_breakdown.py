#!/usr/bin/env python3
"""Where the end-to-end frame time goes (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pathfinder_b200 import api, scenes

def timeit(f, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sum(ts) / len(ts)

for name, (flat, xf), size in [("tiger4k", scenes.tiger(4096), 4096), ("random100k", (scenes.random_paths(100000, 8192, 0x5EED0004), None), 8192)]:
    r = api.CudaRenderer((size, size), background_color=(1, 1, 1, 1))
    scene = api.Scene.from_flat(flat)
    opts = api.BuildOptions(transform=None if xf is None else api.Transform2F(*xf))
    scene.build_and_render(r, opts); r.synchronize()
    host = torch.empty((size, size, 4), dtype=torch.uint8, pin_memory=True)
    def clean(): scene.build_and_render(r, opts); r.synchronize()
    def dirty(): scene.set_view_box(flat.view_box); scene.build_and_render(r, opts); r.synchronize()
    def readback(): r.read_pixels_into(host.data_ptr(), size * 4)
    print(name, "clean frame ms (min, mean):", timeit(clean), "dirty frame:", timeit(dirty), "readback:", timeit(readback),
          "GB/s:", size * size * 4 / timeit(readback)[0] / 1e6)
    r.set_deferred_verification(True)
    def dirty_host():  # host time of a dirty frame when nothing waits for the GPU
        scene.set_view_box(flat.view_box); scene.build_and_render(r, opts)
    print("   deferred verification, dirty frame: host-only ms (min, mean):", timeit(dirty_host, 8))
    print("   stats", {k: v for k, v in r.stats().items() if k in ("host_sync_count", "h2d_bytes", "batch_cache_hits", "cpu_build_time_ns")})
