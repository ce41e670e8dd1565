#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_gather.py : the strip-partitioned frame assembled by the fused
peer-store gather equals the frame assembled by NCCL all-gather and the single-GPU frame, on every rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pathfinder_b200 import api, scenes, partition

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = 2048
flat = scenes.random_paths(6000, size, 77, r_min=8.0, r_max=160.0)
scene = api.Scene.from_flat(flat)
opts = api.BuildOptions()
stream = torch.cuda.current_stream()

def make(peer):
    full = torch.zeros((size, size, 4), dtype=torch.uint8, device="cuda")
    r = api.CudaRenderer((size, size), background_color=(1, 1, 1, 1), device_ordinal=local)
    r.set_stream(stream.cuda_stream)
    r.set_dest_device_pointer(full.data_ptr(), size * 4)
    return r, full

ref_r, ref = make(False)
scene.build_and_render(ref_r, opts)          # full frame on every rank
y0, y1 = partition.strip_rows(size, world, rank)

r1, f1 = make(False)
r1.set_strip(y0, y1)
scene.build_and_render(r1, opts)
dist.all_gather_into_tensor(f1.view(-1), f1[y0 * 16:y1 * 16].reshape(-1))

r2, f2 = make(True)
mine = api.ipc_export(f2.data_ptr())
everyone = [None] * world
dist.all_gather_object(everyone, mine)
peers = [everyone[i] for i in range(world) if i != rank]
r2.set_peer_dests([h for h, _ in peers], [o for _, o in peers])
r2.set_strip(y0, y1)
for _ in range(3):
    scene.build_and_render(r2, opts)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    dist.all_reduce(flag)
torch.cuda.synchronize()
ok = bool(torch.equal(f1, ref)) and bool(torch.equal(f2, ref))
print(f"rank {rank}: nccl==full {bool(torch.equal(f1, ref))}, peer==full {bool(torch.equal(f2, ref))}")
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
r2.set_peer_dests([], [])
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
