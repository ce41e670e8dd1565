#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_gather.py : the strip-partitioned frame assembled inside the library
(PFCudaRendererGatherFrame: ncclAllGather for equal strips, grouped ncclBroadcast for unequal ones) and by the fused
peer-store gather both equal the single-GPU frame, on every rank, byte for byte; repeated frames stay identical
(the next frame's compositing is ordered after the gather in flight)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pathfinder_b200 import api, scenes

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.current_stream()
ok = True


def make(width, height):
    full = torch.zeros((height, width, 4), dtype=torch.uint8, device="cuda")
    r = api.CudaRenderer((width, height), background_color=(1, 1, 1, 1), device_ordinal=local)
    r.set_stream(stream.cuda_stream)
    r.set_dest_device_pointer(full.data_ptr(), width * 4)
    return r, full


# equal strips (128 tile rows) and unequal ones (a 1000-px-high frame: 63 tile rows, the last one partial)
for width, height in [(2048, 2048), (1536, 1000)]:
    flat = scenes.random_paths(4000, max(width, height), 77, r_min=8.0, r_max=160.0)
    scene = api.Scene.from_flat(flat)
    scene.set_view_box((0.0, 0.0, float(width), float(height)))
    opts = api.BuildOptions()
    ref_r, ref = make(width, height)
    scene.build_and_render(ref_r, opts)  # the whole frame on every rank
    ref_r.synchronize()

    # the library's gather
    r1, f1 = make(width, height)
    box = [api.gather_create_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    r1.gather_init(box[0], rank, world)
    for mode, mode_name in ((api.CudaRenderer.GATHER_MODE_FRAME, "frame"), (api.CudaRenderer.GATHER_MODE_TILES, "tiles")):
        r1.gather_set_mode(mode)
        for deferred in (False, True):
            r1.set_deferred_verification(deferred)
            f1.zero_()
            for _ in range(4):
                scene.build_and_render(r1, opts)
                r1.gather_frame()
            r1.synchronize()
            same = bool(torch.equal(f1, ref))
            ok = ok and same
            print(f"rank {rank}: {width}x{height} library gather, {mode_name} mode (deferred={deferred}) == full frame: {same}", flush=True)
    r1.gather_destroy()

    # fused peer stores (IPC-mapped frames)
    r2, f2 = make(width, height)
    mine = api.ipc_export(f2.data_ptr())
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    peers = [everyone[i] for i in range(world) if i != rank]
    r2.set_peer_dests([h for h, _ in peers], [o for _, o in peers])
    r2.set_strip(*api.strip_of_rank((height + 15) // 16, rank, world))
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(3):
        scene.build_and_render(r2, opts)
        dist.all_reduce(flag)
    torch.cuda.synchronize()
    same = bool(torch.equal(f2, ref))
    ok = ok and same
    print(f"rank {rank}: {width}x{height} peer-store gather == full frame: {same}", flush=True)
    dist.barrier()
    r2.set_peer_dests([], [])

t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
