import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pathfinder_b200 import api, scenes
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
items = []
for name, (flat, xf), size in [("tiger", scenes.tiger(4096), 4096), ("random", (scenes.random_paths(100000, 8192, 0x5EED0004), None), 8192)]:
    r = api.CudaRenderer((size, size), background_color=(1, 1, 1, 1))
    r.set_stream(stream.cuda_stream)
    buf = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
    r.set_dest_device_pointer(buf.data_ptr(), size * 4)
    sc = api.Scene.from_flat(flat)
    op = api.BuildOptions(transform=None if xf is None else api.Transform2F(*xf))
    sc.build_and_render(r, op); r.synchronize()
    ref = buf.clone()
    items.append((name, r, sc, op, buf, ref))
for deferred in (False, True):
    for _, r, *_ in items: r.set_deferred_verification(deferred)
    for _ in range(3):
        for _, r, sc, op, *_ in items: sc.build_and_render(r, op)
    torch.cuda.synchronize()
    for _, r, sc, op, buf, ref in items: buf.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); s.record(stream)
    for _ in range(20):
        for _, r, sc, op, *_ in items: sc.build_and_render(r, op)
    e.record(stream); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("deferred", deferred, "events ms/step %.3f host enqueue ms/step %.3f wall ms/step %.3f" % (s.elapsed_time(e) / 20, (t1 - t0) * 50, (t2 - t0) * 50))
    for name, r, sc, op, buf, ref in items:
        print("   ", name, "frame ok:", bool(torch.equal(buf, ref)), {k: v for k, v in r.stats().items() if k in ("reruns", "host_sync_count", "batch_cache_hits", "drawcall_count")})
