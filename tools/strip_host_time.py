#!/usr/bin/env python3
"""Host time of a rank's dirty frame at N ranks, measured on ONE GPU: the renderer owns the strip a middle rank would
own, the host worker pool is capped at the rank's share of the cores, every frame re-uploads the scene (epoch bump)
and nothing waits for the GPU (deferred verification). Prints the host-only time per frame (what bounds the end-to-end
step at 4 and 8 GPUs) and the wall time per frame with one synchronisation at the end.

    PF_HOST_THREADS=2 python tools/strip_host_time.py 8        # a rank's view of random100k@8192 at 8 GPUs
    PF_CUDA_LIB=ab/libpf_base.so ... the same against another build of the library (A/B)
"""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathfinder_b200 import api, scenes  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = 8192
flat = scenes.random_paths(100000, size, 0x5EED0004)
renderer = api.CudaRenderer((size, size), background_color=(1.0, 1.0, 1.0, 1.0))
scene = api.Scene.from_flat(flat)
options = api.BuildOptions()
scene.build_and_render(renderer, options)
renderer.synchronize()
y0, y1 = api.strip_of_rank(size // 16, world // 2, world)
if world > 1:
    renderer.set_strip(y0, y1)
renderer.set_deferred_verification(True)
for _ in range(5):
    scene.set_view_box(flat.view_box)
    scene.build_and_render(renderer, options)
renderer.synchronize()
host = []
frames = 40
t_all = time.perf_counter()
for _ in range(frames):
    t = time.perf_counter()
    scene.set_view_box(flat.view_box)  # epoch bump: segments rebuilt and re-uploaded
    scene.build_and_render(renderer, options)
    host.append((time.perf_counter() - t) * 1e3)
renderer.synchronize()
wall = (time.perf_counter() - t_all) * 1e3 / frames
print("lib=%s threads=%s world=%d rows=[%d,%d): host per dirty frame median %.3f ms (min %.3f), wall per frame %.3f ms"
      % (os.environ.get("PF_CUDA_LIB", "in-tree"), os.environ.get("PF_HOST_THREADS", "all"), world, y0, y1,
         statistics.median(host), min(host), wall))
