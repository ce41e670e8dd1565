#!/usr/bin/env python3
"""bench.py — the driver's benchmark contract for the CUDA backend.

    python bench.py --gpus N --steps K --warmup W            # our arm (N > 1: under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm = CPU tiler on host cores

One "step" renders one frame of each scene of the headline workload (BASELINE.json metric:
"ms/frame + Gsegments/s, tiger@4K & 100k-path@8K"): the Ghostscript tiger at 4096x4096 in its
all-winding and odd-paths-even-odd variants (BASELINE.json configs[1]) and the synthetic 100k
random cubic paths at 8192x8192 (configs[3]). A *segment* is one flattened line segment entering
bin (one process_line_segment call, renderer/src/tiler.rs:177). `value` = segments of all frames of
the step / device time with the scenes resident in HBM; `e2e` = the same through the public API
with the scene re-uploaded from host memory and the frame read back to pinned host memory every
frame. With N > 1 GPUs every frame is partitioned by horizontal tile strips (one rank per GPU) and
assembled on every rank by the library itself (PFCudaRendererGatherFrame: ncclAllGather on the renderer's
gather stream), so scaling is strong (fixed work per step). torch.distributed only boots the ranks, ships
the gather id and reduces the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Gsegments/s"
UNIT = "Gsegments/s"


def workload_scenes(which: str):
    """Returns [(name, FlatScene, transform | None, size)]."""
    from pathfinder_b200 import scenes
    out = []
    if which in ("headline", "tiger4k"):
        flat, xf = scenes.tiger(4096)
        out.append(("tiger@4096/winding", flat, xf, 4096))
        flat, xf = scenes.tiger(4096, even_odd_odd_paths=True)
        out.append(("tiger@4096/even-odd", flat, xf, 4096))
    if which in ("headline", "random100k"):
        out.append(("random100k@8192", scenes.random_paths(100000, 8192, 0x5EED0004), None, 8192))
    if which == "random1m":
        out.append(("random1m@16384", scenes.random_paths(1000000, 16384, 0x5EED0005), None, 16384))
    if which == "text10k":
        # BASELINE.json configs[2]: the glyphs, white and x-scaled by 3, for the 3x-wide render target; device_frames()
        # wraps them like the reference's demo (render target + page rectangle under PatternFilter::Text)
        out.append(("text10k@2048/subpixel", scenes.text_page_subpixel(10000, 2048), None, 2048))
    if which == "smoke":
        flat, xf = scenes.tiger(512)
        out.append(("tiger@512", flat, xf, 512))
    if not out:
        raise SystemExit(f"unknown workload {which!r}")
    return out


WORKLOAD_NAMES = {
    "headline": "tiger@4096 (winding + even-odd variants) + random100k@8192",
    "tiger4k": "tiger@4096 (winding + even-odd variants)",
    "random100k": "random100k@8192",
    "random1m": "random1m@16384",
    "text10k": "text page: 10,000 Roboto glyphs at 12-16 px @2048, subpixel AA (3x-wide render target + text filter: "
               "defringing kernel, gamma LUT), stem darkening",
    "smoke": "tiger@512",
}


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for p in self.rows:
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


STEM_DARKENING_16PX = (0.0121 * 16 * 3, 0.0121 * 1.25 * 16)  # STEM_DARKENING_FACTORS * font size, x in subpixels


def oracle_scene_and_options(flat, xf, dilation=(0.0, 0.0)):
    from oracle import pf_oracle as O
    sc = O.make_scene(points=flat.points, point_flags=flat.point_flags, contour_offsets=flat.contour_offsets,
                      draw_contour_ranges=flat.contour_ranges(), draw_fill_rules=flat.fill_rules,
                      draw_paints=flat.paints, paint_colors=flat.paint_colors, view_box=flat.view_box)
    t = None if xf is None else (xf[0], xf[2], xf[1], xf[3], xf[4], xf[5])
    return sc, O.make_options(transform=t, dilation=dilation)


def cpu_tiler_step(prepared, n_threads: int):
    """One step of the reference arm: the CPU tiler (oracle port of renderer/src/tiler.rs +
    builder.rs, D3D9 level) over every scene of the workload. Returns (seconds, segments)."""
    from oracle import pf_oracle as O
    secs, segs = 0.0, 0
    for sc, opt, n_segments in prepared:
        secs += O.time_build(sc, opt, n_threads)
        segs += n_segments
    return secs, segs


def prepare_cpu(scene_list):
    from oracle import pf_oracle as O
    prepared = []
    for name, flat, xf, _size in scene_list:
        sc, opt = oracle_scene_and_options(flat, xf, STEM_DARKENING_16PX if name.endswith("/subpixel") else (0.0, 0.0))
        b = O.Built(sc, opt, n_threads=max(1, os.cpu_count() or 1))
        prepared.append((sc, opt, b.line_segment_count))
        b.close()
    return prepared


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores. The
    reference is Rust and cannot be built in this image (no rustc/cargo), so this times the oracle
    port with all host threads; kind = "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, os.cpu_count() or 1)
    scene_list = workload_scenes(args.workload)
    prepared = prepare_cpu(scene_list)
    # The port does not always scale to every core (its batch pack is sequential, like the reference's, and on the GPU
    # box it peaked at half the logical cores): the warm-up steps try all the cores and half of them in turn, and the
    # timed steps run on whichever was faster, so that the baseline is the port at its best, not at its widest.
    candidates = [cores] if cores < 4 else [cores, cores // 2]
    tried = {n: float("inf") for n in candidates}
    for w in range(max(args.warmup, len(candidates))):
        n_threads = candidates[w % len(candidates)]
        tried[n_threads] = min(tried[n_threads], cpu_tiler_step(prepared, n_threads)[0])
    threads = min(candidates, key=lambda n: tried[n])
    t_total, seg_total = 0.0, 0
    for _ in range(args.steps):
        s, n = cpu_tiler_step(prepared, threads)
        t_total += s
        seg_total += n
    value = seg_total / t_total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload],
                   "note": "CPU tiler only (scene build: flatten + tile + propagate + pack), as cpu_build_time"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full builds of every scene of the workload",
                         "host_cores": cores,
                         "warmup_ms_per_step_by_threads": {str(n): round(t * 1e3, 3) for n, t in tried.items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def ncu_dram_bytes(world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fill + tile stage (k_tile_solid + k_tile_alpha, one frame of
    random100k@8192) from the committed `ncu --set full` capture (profiles/r*_fill_tile_traffic.json, newest round).
    Only meaningful for the whole frame (N = 1), and only while composite.cu is still the file that was profiled:
    the capture records its SHA-256 and a stale capture reads as null."""
    import glob
    import hashlib
    if world != 1:
        return None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_fill_tile_traffic.json")))
    if not files:
        return None
    rec = json.load(open(files[-1]))
    src = os.path.join(ROOT, "pathfinder_b200", "csrc", "composite.cu")
    if not os.path.exists(src) or hashlib.sha256(open(src, "rb").read()).hexdigest() != rec.get("composite_cu_sha256"):
        return None
    return int(rec["dram_bytes_per_frame"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOAD_NAMES))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frames-in-flight", type=int, default=1, choices=[1, 2],
                    help="renderer instances per scene in the device-resident timed region: with 2, consecutive frames of a "
                         "scene alternate between two renderers on their own streams (double buffering), so the stages of "
                         "frame i + 1 overlap those of frame i; isolated per-frame times and e2e always use one. Measured on a "
                         "B200: 1.003 -> 0.953 ms/step (+5 %): the three scenes of a step already fill the GPU, so 1 stays "
                         "the default")
    ap.add_argument("--no-random1m", action="store_true", help="skip the random1m@16384 sub-record of the headline line")
    ap.add_argument("--no-e2e-double-buffer", action="store_true",
                    help="end to end at one GPU: one renderer + frame buffer per scene instead of two alternating ones")
    ap.add_argument("--gather", default="tiles", choices=["tiles", "nccl", "peer"],
                    help="N > 1, how the frame is assembled on every rank. 'tiles' (default) = PFCudaRendererGatherFrame in "
                         "tile mode: compact exports (4 B per single-colour tile, 1 KB per other tile) pulled from the peers' "
                         "memory over NVLink after an NCCL barrier; 'nccl' = the same call in frame mode (ncclAllGather of the "
                         "finished strips); 'peer' = the fill+tile kernels store every tile into all peers' frames "
                         "(IPC-mapped buffers) + one barrier. All overlap the next frame's stages.")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    # One process per GPU: the ranks share the host's cores for their scene / batch builds.
    os.environ.setdefault("PF_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // max(1, int(os.environ.get("WORLD_SIZE", "1"))))))

    import torch
    from pathfinder_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the hot path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # Explicit streams: `stream` carries the timing events and barriers; every frame of the step (they are
    # independent scenes with their own renderers) renders on a stream of its own, forked from and joined to
    # `stream` around the timed region, so the GPU can overlap the latency-bound stages of one scene with the
    # throughput-bound stages of another. (The legacy default stream would not do: a renderer given stream 0 falls
    # back to a private non-blocking stream that events on the default stream do not see.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    scene_list = workload_scenes(args.workload)

    # One renderer per scene size; every rank owns a strip of tile rows and renders straight into
    # its slot of the gathered frame.
    class Frame:
        pass

    def shared_frame_path(index, size):
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize > 2 * size * size * 4 else "/tmp"
        return os.path.join(base, "pf_bench_%s_%d_%d.frame" % (os.environ.get("MASTER_PORT", "0"), index, size))

    frames = []

    def make_frame(name, flat, xf, size, index, device_only=False):
        f = Frame()
        f.name, f.flat, f.size = name, flat, size
        f.y0, f.y1 = api.strip_of_rank((size + 15) // 16, rank, world)
        f.full = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
        f.renderer = api.CudaRenderer((size, size), background_color=(1.0, 1.0, 1.0, 1.0), device_ordinal=local_rank)
        f.stream = torch.cuda.Stream()
        f.renderer.set_stream(f.stream.cuda_stream)
        f.renderer.set_dest_device_pointer(f.full.data_ptr(), size * 4)
        if name.endswith("/subpixel"):
            from pathfinder_b200 import scenes as _scenes
            f.scene = _scenes.subpixel_scene(flat, size)
            f.options = api.BuildOptions(dilation=STEM_DARKENING_16PX)
        else:
            f.scene = api.Scene.from_flat(flat)
            f.options = api.BuildOptions(transform=None if xf is None else api.Transform2F(*xf))
        # The metric counts the flattened segments of the WHOLE frame (fixed work, independent of
        # N): one untimed full-frame render gives the count before the strip is set.
        f.scene.build_and_render(f.renderer, f.options)
        f.full_stats = f.renderer.stats()
        # Frames are enqueued back to back: the check that no stage overflowed its bound happens at
        # the next use of the renderer instead of stalling the host at the end of every batch.
        f.renderer.set_deferred_verification(True)
        if world > 1:
            if args.gather in ("nccl", "tiles"):
                # The library assembles the frame: rank 0 creates the group id, torch.distributed only ships it.
                box = [api.gather_create_id() if rank == 0 else None]
                dist.broadcast_object_list(box, src=0)
                f.renderer.gather_init(box[0], rank, world)  # also sets this rank's strip
                if args.gather == "nccl":
                    f.renderer.gather_set_mode(api.CudaRenderer.GATHER_MODE_FRAME)
                else:  # tiles: refuse to report a frame-mode number under the tile-mode label
                    f.renderer.gather_set_mode(api.CudaRenderer.GATHER_MODE_TILES)
            else:
                f.renderer.set_strip(f.y0, f.y1)
            f.strip_view = f.full[f.y0 * 16:min(f.y1 * 16, size)]
        f.host = torch.empty((size, size, 4), dtype=torch.uint8, pin_memory=True) if (rank == 0 and world == 1 and not device_only) else None
        f.host_shared = None
        if world > 1 and not device_only:
            # End to end at N > 1 the frame is assembled in HOST memory: one shared, page-locked frame that every
            # rank's GPU writes its strip into over its own PCIe link (no device-side gather on this path).
            f.shm_path = shared_frame_path(index, size)
            if rank == 0:
                with open(f.shm_path, "wb") as fh:
                    fh.truncate(size * size * 4)
            dist.barrier()
            f.host_shared = torch.from_file(f.shm_path, shared=True, size=size * size * 4, dtype=torch.uint8).view(size, size, 4)
            rc = torch.cuda.cudart().cudaHostRegister(f.host_shared.data_ptr(), size * size * 4, 0)
            f.host_registered = int(rc[0] if isinstance(rc, tuple) else rc) == 0  # else: pageable copies (slower)
        f.copied = None
        f.peer = False
        if world > 1 and args.gather == "peer":
            # Exchange CUDA IPC handles of the frame buffers; every rank then writes its strip into all copies.
            mine = api.ipc_export(f.full.data_ptr())
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            peers = [everyone[i] for i in range(world) if i != rank]
            f.renderer.set_peer_dests([h for h, _ in peers], [o for _, o in peers])
            f.peer = True
        return f

    for name, flat, xf, size in scene_list:
        frames.append(make_frame(name, flat, xf, size, len(frames)))
    # Frames in flight: a second renderer (own stream, own stage buffers, own frame buffer, own gather group) per
    # scene; steps alternate between the two. Every step still renders every scene once, into a complete frame.
    for f in frames:
        f.alt = None
    if args.frames_in_flight == 2:
        for i, (name, flat, xf, size) in enumerate(scene_list):
            frames[i].alt = make_frame(name, flat, xf, size, len(scene_list) + i, device_only=True)
            frames[i].alt.alt = None
    instances = [g for f in frames for g in ((f, f.alt) if f.alt is not None else (f,))]

    def instance_of(f, step_index):
        return f.alt if (f.alt is not None and step_index % 2 == 1) else f

    copy_stream = torch.cuda.Stream()
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")

    def render_frame(f, e2e: bool):
        with torch.cuda.stream(f.stream):  # NCCL orders its work after the current stream
            render_frame_on_stream(f, e2e)

    def render_frame_on_stream(f, e2e: bool):
        if e2e:
            f.scene.set_view_box(f.flat.view_box)  # bumps the epoch: the scene is re-uploaded from host memory
            if f.copied is not None:
                f.stream.wait_event(f.copied)  # the previous read-back of this frame buffer must be done
        f.scene.build_and_render(f.renderer, f.options)
        if dist is not None and not e2e:
            if f.peer:
                dist.all_reduce(flag)  # barrier on the stream: every rank's strip has landed in every frame copy
            else:
                # Asynchronous: on the renderer's gather stream after this frame's kernels, while the next frame's
                # stages run; the library orders the next compositing of this buffer after it.
                f.renderer.gather_frame()
        if e2e and dist is not None:
            # Every rank reads its own strip back into the shared host frame.
            done = torch.cuda.Event()
            done.record(f.stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                f.host_shared[f.y0 * 16:f.y1 * 16].copy_(f.strip_view, non_blocking=True)
                f.copied = torch.cuda.Event()
                f.copied.record(copy_stream)
        elif e2e:
            # Device -> pinned host read-back of the assembled frame on a copy stream, so it overlaps
            # the host-side build and the rendering of the next frame.
            done = torch.cuda.Event()
            done.record(f.stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                f.host.copy_(f.full, non_blocking=True)
                f.copied = torch.cuda.Event()
                f.copied.record(copy_stream)

    def fork():  # the frames' streams start after everything already on `stream` (the start event)
        for f in instances + [f.e2e_alt for f in frames if getattr(f, "e2e_alt", None) is not None]:
            f.stream.wait_stream(stream)

    def join():  # ... and `stream` (the end event) waits for all of them, frame assembly included
        for f in instances + [f.e2e_alt for f in frames if getattr(f, "e2e_alt", None) is not None]:
            if dist is not None and not f.peer:
                f.renderer.gather_wait()  # the frame's stream waits for its last gather
            stream.wait_stream(f.stream)

    e2e_counter = [0]

    def step(e2e: bool):
        # End to end, one GPU: two instances of every scene's renderer + frame buffers alternate, so the read-back of
        # frame i (PCIe, the long pole) runs while frame i + 1 is built and rendered into the other buffer.
        k = e2e_counter[0]
        e2e_counter[0] += 1
        for f in frames:
            g = f.e2e_alt if (e2e and getattr(f, "e2e_alt", None) is not None and k % 2 == 1) else f
            render_frame(g, e2e)

    def timed(e2e: bool, steps: int):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        start.record(stream)
        fork()
        for _ in range(steps):
            step(e2e)
        join()
        end.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev = start.elapsed_time(end) / 1e3
        # e2e includes host work (scene build, copies): wall clock around the synchronised region;
        # device-resident value: CUDA events on the launching stream.
        t = torch.tensor([wall if e2e else dev], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for w in range(max(args.warmup, 2 * args.frames_in_flight)):
        for f in frames:
            render_frame(instance_of(f, w), False)
    barrier()

    # Per-frame stats after warm-up.
    seg_per_step, launches_per_step = 0, 0
    per_scene = {}
    for f in frames:
        s = f.renderer.stats()
        seg_per_step += f.full_stats["line_segment_count"]
        launches_per_step += s["drawcall_count"]
        per_scene[f.name] = s

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for f in instances:
        f.renderer.set_timing_enabled(True)
    stage_acc = {f.name: {} for f in frames}

    # Timed region: exactly K steps, device-resident scenes.
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    fork()
    for k in range(args.steps):
        for f in frames:
            render_frame(instance_of(f, k), False)  # no host wait in here: verification is deferred, stage events are read afterwards
    join()
    end.record(stream)
    barrier()
    for f in frames:
        totals, batches = f.renderer.accumulated_times()
        if f.alt is not None:  # the scene's frames were shared between its two renderers
            more, more_batches = f.alt.renderer.accumulated_times()
            totals = {k: v + more.get(k, 0.0) for k, v in totals.items()}
            batches += more_batches
            f.alt.renderer.set_timing_enabled(False)
        assert batches and batches % args.steps == 0, (batches, args.steps)  # (several draw batches per frame: render targets)
        stage_acc[f.name] = totals
    dev_s = start.elapsed_time(end) / 1e3
    t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # "ms/frame" proper: every scene alone on the GPU, K frames back to back on its stream (no other scene's
    # kernels in the shadow), frame assembly included; CUDA events around the K frames, max over ranks. The stage
    # times of the same frames feed the roofline: a kernel timed alone, not under three-way overlap.
    isolated_ms, isolated_stage = {}, {}

    def isolated_run(f, steps):
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        f.stream.wait_stream(stream)
        for _ in range(steps):
            render_frame(f, False)
        if dist is not None and not f.peer:
            f.renderer.gather_wait()
        stream.wait_stream(f.stream)
        s1.record(stream)
        barrier()
        ms = torch.tensor([s0.elapsed_time(s1) / steps], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for f in frames:
        f.renderer.set_timing_enabled(False)
        isolated_ms[f.name] = isolated_run(f, args.steps)  # without the stage events (eight event records per batch)
        f.renderer.set_timing_enabled(True)   # resets the accumulated stage times
        isolated_run(f, args.steps)
        totals, batches = f.renderer.accumulated_times()
        assert batches and batches % args.steps == 0, (batches, args.steps)
        isolated_stage[f.name] = {k: v / args.steps for k, v in totals.items()}
        f.renderer.set_timing_enabled(False)

    # End to end: scene upload from host memory + render + read-back of the frame, every frame.
    e2e_steps = max(4, min(args.steps, 10))
    e2e_in_flight = 1
    if world == 1 and not args.no_e2e_double_buffer:
        for i, (name, flat, xf, size) in enumerate(scene_list):
            frames[i].e2e_alt = make_frame(name, flat, xf, size, 2 * len(scene_list) + i)
            frames[i].e2e_alt.alt = None
        e2e_in_flight = 2
    for _ in range(2 * e2e_in_flight):
        step(True)
    e2e_s = timed(True, e2e_steps)
    h2d = sum(int(f.flat.points.nbytes + f.flat.n_contours * 8 + len(f.flat.points) * 8) for f in frames)
    d2h = sum(f.size * f.size * 4 for f in frames)

    # BASELINE.json configs[3] (random1m@16384: 1 GiB of frame, the case the strips are for) alone on the GPUs, device
    # resident, same timing as the isolated frames above. A sub-record of the headline line, not the headline value.
    random1m = None
    if args.workload == "headline" and not args.no_random1m:
        (name, flat, xf, size), = workload_scenes("random1m")
        f = make_frame(name, flat, xf, size, len(frames), device_only=True)
        r1m_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            render_frame(f, False)
        ms = isolated_run(f, r1m_steps)
        segs = f.full_stats["line_segment_count"]
        random1m = {"workload": WORKLOAD_NAMES["random1m"], "ms_per_step": ms, "steps": r1m_steps, "segments_per_step": segs,
                    "value": segs / (ms * 1e-3) / 1e9, "unit": UNIT,
                    "timing": "CUDA events around the steps, frame assembly included, max over ranks; scene resident in HBM"}
        if dist is not None and args.gather in ("nccl", "tiles"):
            f.renderer.gather_destroy()
        del f
        torch.cuda.empty_cache()

    e2e_assembly = ("device frame -> pinned host frame; two renderer instances + frame buffers per scene alternate, so a "
                    "frame's read-back overlaps the next frame's build and render" if e2e_in_flight == 2
                    else "device frame -> pinned host frame")
    if dist is not None:
        e2e_assembly = ("every rank copies its strip into one shared %s host frame over its own PCIe link"
                        % ("page-locked" if all(f.host_registered for f in frames) else "pageable"))
        torch.cuda.synchronize()
        for f in frames:
            if f.host_registered:
                torch.cuda.cudart().cudaHostUnregister(f.host_shared.data_ptr())
        dist.barrier()
        if rank == 0:
            for f in frames:
                try:
                    os.unlink(f.shm_path)
                except OSError:
                    pass
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = seg_per_step * args.steps / dev_s / 1e9
    inputs_per_step = sum(int(f.full_stats.get("input_segment_count", 0)) for f in frames)
    fills_per_step = sum(int(f.full_stats.get("fill_count", 0)) for f in frames)
    e2e_value = seg_per_step * e2e_steps / e2e_s / 1e9

    # Roofline of the dominant stage (fill + tile: k_tile_solid + k_tile_alpha, launched back to back) on the largest
    # scene. Algorithmic bytes per SURVEY.md §8(d) "fused fill+composite": 12 B per visible fill (the reference's
    # Fill record) + 16 B per tile-list entry (TileD3D11) + 4 B per pixel written. (What this implementation
    # actually moves is 8 B per fill and 32 B per entry: `implementation_bytes_per_launch`.)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    big = max(frames, key=lambda f: f.size)
    bs = per_scene[big.name]
    rows = (big.y1 - big.y0) * 16
    # Only the fills of tiles that survive the z-cull are ever stored or read.
    algo_bytes = 12 * bs["visible_fill_count"] + 16 * bs["tile_list_entry_count"] + 4 * big.size * rows
    impl_bytes = 8 * bs["visible_fill_count"] + 32 * bs["tile_list_entry_count"] + 4 * big.size * rows
    kernel_ms = isolated_stage[big.name]["fill_tile_ms"]
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "fill + tile: k_tile_solid + k_tile_alpha (one stage, two launches)", "scene": big.name,
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_dram_bytes(world),
                "algorithmic_bytes_per_launch": algo_bytes, "implementation_bytes_per_launch": impl_bytes,
                "kernel_ms": kernel_ms, "kernel_ms_in_overlapped_step": stage_acc[big.name]["fill_tile_ms"] / args.steps,
                "timing": "CUDA events around the stage on the renderer's stream, the scene rendering alone (isolated phase)",
                "peak_source": peak_src}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cores = max(1, os.cpu_count() or 1)
        prepared = prepare_cpu(scene_list)
        cpu_tiler_step(prepared, cores)
        reps, t_cpu, seg_cpu = 0, 0.0, 0
        t_begin = time.perf_counter()
        while reps < 3 or (time.perf_counter() - t_begin < 10.0 and reps < 50):
            s, n = cpu_tiler_step(prepared, cores)
            t_cpu += s
            seg_cpu += n
            reps += 1
        s1, n1 = cpu_tiler_step(prepared, 1)
        # how the port scales with threads (one build each below all cores): how far the best figure is from
        # cores x the single-thread figure says how much a better-parallelised tiler could gain
        scaling = {"1": n1 / s1 / 1e9}
        for nt in (2, 4, 8, 16, 32, 64):
            if nt < cores:
                st_, nn = cpu_tiler_step(prepared, nt)
                scaling[str(nt)] = nn / st_ / 1e9
        scaling[str(cores)] = seg_cpu / t_cpu / 1e9
        # (the baseline is the port at its best thread count, as in the reference arm: on the GPU box it peaks below
        # the number of logical cores)
        best_threads = max(scaling, key=lambda k: scaling[k])
        best_reps = reps
        if int(best_threads) != cores:  # measured on one build so far: give it the sample the all-core figure had
            best_reps, t_best, seg_best = 0, 0.0, 0
            t_begin = time.perf_counter()
            while best_reps < 3 or (time.perf_counter() - t_begin < 5.0 and best_reps < 50):
                s, n = cpu_tiler_step(prepared, int(best_threads))
                t_best += s
                seg_best += n
                best_reps += 1
            scaling[best_threads] = seg_best / t_best / 1e9
            if scaling[str(cores)] > scaling[best_threads]:
                best_threads, best_reps = str(cores), reps
        cpu_baseline = {"value": scaling[best_threads], "unit": UNIT, "cores": int(best_threads), "kind": "port",
                        "host_cores": cores, "thread_scaling": scaling,
                        "sample": f"{best_reps} full CPU-tiler builds of every scene of the workload (scene build only: "
                                  "flatten + tile + propagate + pack, the reference's cpu_build_time)",
                        "single_thread_value": n1 / s1 / 1e9}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "frames_per_step": len(frames),
                   "segments_per_step": seg_per_step,
                   # the same step counted in input segments (SegmentIndicesD3D11) and in fills
                   "input_segments_per_step": inputs_per_step, "fills_per_step": fills_per_step,
                   "input_gsegments_per_s": inputs_per_step * args.steps / dev_s / 1e9,
                   "gfills_per_s": fills_per_step * args.steps / dev_s / 1e9,
                   "parallelism": f"tile-strip x{world}" + ((" + fused peer-store gather (NVLink P2P) + barrier" if args.gather == "peer"
                                                             else (" + compact tile exports pushed to the peers over NVLink by the library (PFCudaRendererGatherFrame, tile mode), overlapped with the next frame" if args.gather == "tiles"
                                                                   else " + ncclAllGather issued by the library (PFCudaRendererGatherFrame, frame mode), overlapped with the next frame")) if world > 1 else ""),
                   "l2": "inputs larger than L2: a step touches > 1 GB of stage buffers and frames",
                   "frames_in_flight": args.frames_in_flight,
                   "streams": "the frames of a step are independent scenes and render concurrently on one CUDA stream each "
                              "(forked from / joined to the timing stream); consecutive frames of a scene alternate between "
                              "%d renderer instance(s) with their own streams, stage buffers and frame buffers; no host wait "
                              "inside the timed region (deferred verification); ms_per_frame / stage_ms are per-stream event "
                              "times and overlap; ms_per_frame_isolated = one scene, one renderer, alone on the GPU" % args.frames_in_flight,
                   "random1m": random1m,
                   "ms_per_frame_isolated": isolated_ms,
                   "stage_ms_isolated": isolated_stage,
                   "ms_per_frame": {f.name: stage_acc[f.name]["total_ms"] / args.steps for f in frames},
                   "stage_ms": {f.name: {k: v / args.steps for k, v in stage_acc[f.name].items()} for f in frames}},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "assembly": e2e_assembly,
                "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps, "frames_in_flight": e2e_in_flight},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
