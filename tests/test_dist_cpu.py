"""CPU, world_size 2 over gloo: the multi-GPU sharding logic - stream chunks with
halo frames cover every pair exactly once, window pairs are partitioned, and the feature
all_gather reassembles every keyframe's record on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200slam import sharding


def test_stream_chunks_partition_all_pairs():
    for n_frames in (2, 7, 500, 10000):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.stream_chunk(n_frames, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(1, n_frames)), (n_frames, world)
    lo, hi = sharding.stream_chunk(10000, 3, 8)
    assert hi - lo in (1249, 1250)


def test_window_pairs_partition():
    pairs = sharding.window_pairs(16)
    assert len(pairs) == 120
    for world in (1, 2, 4, 8):
        got = sorted(p for r in range(world) for p in sharding.shard_pairs(pairs, r, world))
        assert got == sorted(pairs)
        assert max(len(sharding.shard_pairs(pairs, r, world)) for r in range(world)) == -(-120 // world)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_kf, max_kp = 6, 32
        mine = sharding.frames_of_rank(n_kf, rank, world)
        g = torch.Generator().manual_seed(0)
        all_k = torch.rand(n_kf, max_kp, 2, generator=g); all_d = torch.rand(n_kf, max_kp, 128, generator=g)
        all_c = torch.randint(1, max_kp, (n_kf,), generator=g, dtype=torch.int32)
        slots = -(-n_kf // world)
        rec = sharding.WindowRecord(slots, max_kp, "cpu")
        for s, f in enumerate(mine):
            rec.put(s, all_k[f], all_d[f], all_c[f:f + 1])
        gk, gd, gc = sharding.gather_window_records(rec, world)       # ONE all_gather_into_tensor of the flat record
        table = sharding.global_frame_table(n_kf, world)
        ok = gk.shape == (world * slots * max_kp, 2) and gd.shape == (world * slots * max_kp, 128) and gc.dtype == torch.int32
        for f in range(n_kf):
            owner, slot = table[f]
            row = (owner * slots + slot) * max_kp
            ok &= torch.equal(gk[row:row + max_kp], all_k[f]) and torch.equal(gd[row:row + max_kp], all_d[f]) and int(gc[owner * slots + slot]) == int(all_c[f])
        pairs = sharding.shard_pairs(sharding.window_pairs(n_kf), rank, world)
        cnt = torch.tensor([len(pairs)])
        dist.all_reduce(cnt)
        ok &= int(cnt) == 15
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok &= float(t) == float(world)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_reduce():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
