"""N > 1 host logic on CPU: world_size-2 gloo run of the strip partition + all-gather plumbing that
bench.py uses, with the oracle standing in for the per-rank renderer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pathfinder_b200 import area_lut, partition, scenes
from tests import helpers as H

SIZE = 256


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat, xf = scenes.tiger(SIZE)
        y0, y1 = partition.strip_rows(SIZE, world, rank)
        built = H.oracle_build(flat, xf, strip=(y0, y1))
        img = built.render(area_lut.generate(), SIZE, SIZE, background=(1, 1, 1, 1))
        p0, p1 = partition.strip_pixel_rows(SIZE, world, rank)
        # the library's gather for unequal strips: one broadcast per rank, in place (grouped ncclBroadcast)
        full = torch.zeros((SIZE, SIZE, 4), dtype=torch.uint8)
        full[p0:p1] = torch.from_numpy(np.ascontiguousarray(img[p0:p1]))
        for g in range(world):
            g0, g1 = partition.strip_pixel_rows(SIZE, world, g)
            part = full[g0:g1].contiguous()
            dist.broadcast(part, src=g)
            full[g0:g1] = part
        counts = torch.tensor([len(built.fills), len(built.tiles)], dtype=torch.int64)
        dist.all_reduce(counts)
        if rank == 0:
            np.savez(out_path, frame=full.numpy(), counts=counts.numpy())
    finally:
        dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("world", [2, 3])
def test_strip_render_and_gather(tmp_path, world):
    """world = 3: 16 tile rows split 5 / 5 / 6 — the unequal strips the grouped-broadcast path assembles."""
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    z = np.load(out)
    flat, xf = scenes.tiger(SIZE)
    full = H.oracle_build(flat, xf)
    ref = full.render(area_lut.generate(), SIZE, SIZE, background=(1, 1, 1, 1))
    assert np.array_equal(z["frame"], ref)
    assert z["counts"].tolist() == [len(full.fills), len(full.tiles)]


def test_strip_rows():
    assert partition.strip_rows(8192, 8, 3) == (192, 256)
    assert partition.strip_pixel_rows(4096, 2, 1) == (2048, 4096)
    assert [partition.strip_rows(256, 3, g) for g in range(3)] == [(0, 5), (5, 10), (10, 16)]
    assert partition.strip_pixel_rows(100, 2, 1) == (48, 100)  # 7 tile rows, the last one partial
    with pytest.raises(ValueError):
        partition.strip_rows(16, 3, 0)
    # the host restatement agrees with the library's PFCudaStripOfRank
    from pathfinder_b200 import api
    for rows_px, world in [(8192, 8), (4096, 3), (1000, 7), (16384, 5)]:
        for g in range(world):
            assert api.strip_of_rank(partition.tile_rows(rows_px), g, world) == partition.strip_rows(rows_px, world, g)


def _builder_worker(rank, world, port, out_path):
    """Each rank runs the PRODUCT's strip-aware host build (PFSceneBuildForStrip through the Python mirror) for the
    strip PFCudaStripOfRank gives it, and the ranks exchange what they kept."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pathfinder_b200 import api
        from tests.test_scene_host import collect_strip
        flat = scenes.random_paths(3000, 1024, 5, r_min=4.0, r_max=60.0)
        y0, y1 = api.strip_of_rank(1024 // 16, rank, world)
        got = collect_strip(api.Scene.from_flat(flat), api.BuildOptions(), (y0, y1))
        draw = next(r for r in got if r["kind"] == "DrawTilesD3D11")
        up = next(r for r in got if r["kind"] == "UploadSceneD3D11")
        kept = torch.zeros(flat.n_paths, dtype=torch.int64)
        kept[torch.tensor(draw["global_path_ids"], dtype=torch.int64)] = 1
        rows_ok = all(r[1] < y1 and r[3] > y0 for r in draw["rects"])
        stats = torch.tensor([draw["path_count"], draw["segment_count"], up["index_count"], int(rows_ok)], dtype=torch.int64)
        all_stats = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(all_stats, stats)
        dist.all_reduce(kept)  # how many ranks kept each path
        if rank == 0:
            np.savez(out_path, kept=kept.numpy(), stats=torch.stack(all_stats).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_strip_aware_host_build_across_ranks(tmp_path, world):
    """The library's host builder under the N > 1 launch: every path of the whole build is kept by at least one rank
    (exactly the ranks whose rows its tile rect reaches), paths outside the view box by none, and a rank uploads only
    the segments of its own batch."""
    from pathfinder_b200 import api
    from tests.test_scene_host import collect_strip
    out = str(tmp_path / "built.npz")
    mp.spawn(_builder_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    z = np.load(out)
    flat = scenes.random_paths(3000, 1024, 5, r_min=4.0, r_max=60.0)
    whole = next(r for r in collect_strip(api.Scene.from_flat(flat), api.BuildOptions(), None) if r["kind"] == "DrawTilesD3D11")
    want = np.zeros(flat.n_paths, dtype=np.int64)
    strips = [api.strip_of_rank(1024 // 16, g, world) for g in range(world)]
    for pid, rect in zip(whole["global_path_ids"], whole["rects"]):
        want[pid] = sum(1 for y0, y1 in strips if rect[1] < y1 and rect[3] > y0)
    assert np.array_equal(z["kept"], want)
    assert (want[whole["global_path_ids"]] >= 1).all()
    stats = z["stats"]
    assert stats[:, 3].all()                       # every kept rect reaches its rank's rows
    assert (stats[:, 1] == stats[:, 2]).all()      # a rank uploads exactly its batch's segments
    assert stats[:, 0].sum() == want.sum() and stats[:, 1].sum() >= whole["segment_count"]
