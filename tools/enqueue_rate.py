#!/usr/bin/env python3
"""Is a small frame bound by the host's launch rate? For a scene and a strip of 1 / `parts` of its tile rows (what a
rank renders at `parts` GPUs; no frame assembly here), K frames are enqueued back to back with deferred verification:
prints the host time per frame spent enqueueing, the GPU time per frame (wall clock to the final synchronize), and the
per-stage CUDA-event times of the same frames. Run on the GPU box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathfinder_b200 import api, scenes


def run(name, flat, xf, size, parts, frames=200):
    r = api.CudaRenderer((size, size), background_color=(1, 1, 1, 1))
    rows = size // 16
    if parts > 1:
        y0, y1 = api.strip_of_rank(rows, parts // 2, parts)
        r.set_strip(y0, y1)
    scene = api.Scene.from_flat(flat)
    opts = api.BuildOptions(transform=None if xf is None else api.Transform2F(*xf))
    for _ in range(3):
        scene.build_and_render(r, opts)
    r.synchronize()
    r.set_deferred_verification(True)
    out = {"scene": name, "parts": parts}
    for timing in (False, True):
        r.set_timing_enabled(timing)
        scene.build_and_render(r, opts)
        r.synchronize()
        t0 = time.perf_counter()
        for _ in range(frames):
            scene.build_and_render(r, opts)
        t1 = time.perf_counter()
        r.synchronize()
        t2 = time.perf_counter()
        key = "timed" if timing else "plain"
        out[key] = {"host_enqueue_ms_per_frame": round((t1 - t0) * 1e3 / frames, 4),
                    "wall_ms_per_frame": round((t2 - t0) * 1e3 / frames, 4)}
        if timing:
            totals, n = r.accumulated_times()
            out["stage_ms"] = {k: round(v / max(n, 1), 4) for k, v in totals.items()}
    print(json.dumps(out))
    r.close()


if __name__ == "__main__":
    tiger, txf = scenes.tiger(4096)
    rnd = scenes.random_paths(100000, 8192, 0x5EED0004)
    for parts in (1, 2, 8):
        run("tiger@4096", tiger, txf, 4096, parts)
        run("random100k@8192", rnd, None, 8192, parts)
