"""Multi-rank host logic on CPU: world_size-2 gloo process group (frame sharding, no data-path collective)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out):
    sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
    from w2x import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_for_rank(n_frames, rank, world)
    # "render": payload = deterministic function of the frame index (the real engines are per-GPU and independent)
    results = [(f, f * f + 1) for f in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    ordered = sharding.merge_in_order(gathered)
    ms = sharding.max_over_ranks_ms(10.0 + 5.0 * rank)
    if rank == 0:
        out.put((ordered, ms))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_world2_gloo():
    world, n_frames = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    ordered, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ordered == [f * f + 1 for f in range(n_frames)]
    assert ms == 15.0  # max over ranks


def test_sharding_helpers():
    sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
    from w2x import sharding
    assert sharding.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((sharding.frames_for_rank(10, r, 4) for r in range(4)), [])) == list(range(10))
    bands = [sharding.band_rows(6, r, 4) for r in range(4)]
    assert bands == [(0, 2), (2, 4), (4, 5), (5, 6)]
    with pytest.raises(ValueError):
        sharding.merge_in_order([[(0, "a")], [(2, "b")]])
