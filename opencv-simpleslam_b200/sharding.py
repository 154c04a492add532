"""Work sharding across GPUs.  Frames and frame pairs are independent units, so
the data path needs no collective: every rank works on its own contiguous chunk of the stream
(plus one halo frame it re-extracts itself) or on its own slice of the keyframe-window pairs.
Only the optional gather of per-frame features for cross-GPU window pairs communicates
(`gather_window_features`, one all_gather of (count, kpts, desc) records)."""
from __future__ import annotations

from typing import List, Tuple

import torch


def stream_chunk(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous chunk [lo, hi) of pair indices t (pair t = frames (t-1, t), t in 1..n_frames-1)
    owned by `rank`; the rank extracts frames lo-1 .. hi-1 (frame lo-1 is its halo)."""
    n_pairs = max(n_frames - 1, 0)
    base, rem = divmod(n_pairs, world)
    lo = 1 + rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def window_pairs(n_keyframes: int) -> List[Tuple[int, int]]:
    """All unordered keyframe pairs (i < j) of a window (BASELINE config 3: 16 keyframes -> 120 pairs)."""
    return [(i, j) for i in range(n_keyframes) for j in range(i + 1, n_keyframes)]


def shard_pairs(pairs: List[Tuple[int, int]], rank: int, world: int) -> List[Tuple[int, int]]:
    """Round-robin ownership: pair p on rank p mod world."""
    return pairs[rank::world]


def frames_of_rank(n_keyframes: int, rank: int, world: int) -> List[int]:
    """Keyframe f is extracted on rank f mod world."""
    return list(range(rank, n_keyframes, world))


def gather_window_features(kpts: torch.Tensor, desc: torch.Tensor, counts: torch.Tensor, group=None):
    """all_gather of this rank's extracted keyframes so every rank can match any window pair.
    kpts [F_local, max_kp, 2], desc [F_local, max_kp, 128], counts [F_local] (int32), same
    F_local on every rank (pad with count 0).  Returns lists indexed by source rank.
    NCCL on GPUs (NVLink/NVSwitch), gloo in the CPU tests."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out_k = [torch.empty_like(kpts) for _ in range(world)]
    out_d = [torch.empty_like(desc) for _ in range(world)]
    out_c = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(out_k, kpts.contiguous(), group=group)
    dist.all_gather(out_d, desc.contiguous(), group=group)
    dist.all_gather(out_c, counts.contiguous(), group=group)
    return out_k, out_d, out_c


def global_frame_table(n_keyframes: int, world: int):
    """frame f -> (owner rank, local slot) under frames_of_rank's layout."""
    return {f: (f % world, f // world) for f in range(n_keyframes)}
