#!/usr/bin/env python3
"""Per-stage CUDA-event times of the pipeline on the benchmark scenes (run on the GPU box)."""
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pathfinder_b200 import api, scenes


def run(name, flat, xf, size, frames=5):
    r = api.CudaRenderer((size, size), background_color=(1, 1, 1, 1))
    r.set_timing_enabled(True)
    scene = api.Scene.from_flat(flat)
    opts = api.BuildOptions(transform=None if xf is None else api.Transform2F(*xf))
    best = None
    for i in range(frames):
        t0 = time.perf_counter()
        scene.build_and_render(r, opts)
        r.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        t = r.times()
        t["wall_ms"] = wall
        if best is None or t["total_ms"] < best["total_ms"]:
            best = t
    s = r.stats()
    print(json.dumps({"scene": name, "times": {k: round(v, 4) for k, v in best.items()}, "stats": s}))
    r.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiger4k", "random100k"]
    if "tiger4k" in which:
        flat, xf = scenes.tiger(4096)
        run("tiger@4096", flat, xf, 4096)
    if "random100k" in which:
        flat = scenes.random_paths(100000, 8192, 0x5EED0004)
        run("random100k@8192", flat, None, 8192)
    if "random1m" in which:
        flat = scenes.random_paths(1000000, 16384, 0x5EED0005)
        run("random1m@16384", flat, None, 16384, frames=3)
