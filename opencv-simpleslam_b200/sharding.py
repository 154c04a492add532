"""Work sharding across GPUs.  Frames and frame pairs are independent units, so
the data path needs no collective: every rank works on its own contiguous chunk of the stream
(plus one halo frame it re-extracts itself) or on its own slice of the keyframe-window pairs.
Only the gather of per-frame features for cross-GPU window pairs communicates
(`gather_window_records`: one all_gather_into_tensor of a flat (kpts | desc | count) record buffer)."""
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


class WindowRecord:
    """This rank's keyframe records as ONE flat f32 buffer, so that sharing a window is a single collective:
    [ keypoints slots*max_kp*2 | descriptors slots*max_kp*128 | counts slots ]   (counts stored as exact floats)."""

    def __init__(self, slots: int, max_kp: int, device):
        self.slots, self.max_kp = int(slots), int(max_kp)
        self.nk, self.nd = self.slots * self.max_kp * 2, self.slots * self.max_kp * 128
        self.flat = torch.zeros(self.nk + self.nd + self.slots, dtype=torch.float32, device=device)

    @property
    def kpts(self):
        return self.flat[:self.nk].view(self.slots, self.max_kp, 2)

    @property
    def desc(self):
        return self.flat[self.nk:self.nk + self.nd].view(self.slots, self.max_kp, 128)

    @property
    def counts(self):
        return self.flat[self.nk + self.nd:]

    def put(self, slot: int, kpts: torch.Tensor, desc: torch.Tensor, n: torch.Tensor):
        """Device-side copy of one extraction (handle-owned buffers [max_kp, .], count on the device): no host sync."""
        self.kpts[slot].copy_(kpts[:self.max_kp]); self.desc[slot].copy_(desc[:self.max_kp])
        self.counts[slot:slot + 1].copy_(n.to(torch.float32))


def gather_window_records(rec: WindowRecord, world: int, group=None):
    """ONE `all_gather_into_tensor` of the flat records (17 MB for 16 keyframes of 2048 points; NCCL on GPUs, gloo in the
    CPU tests), then the slot-major views every rank matches from:
      kpts [world*slots*max_kp, 2], desc [world*slots*max_kp, 128] (frame of rank r, slot s at row (r*slots+s)*max_kp),
      counts int32 [world*slots]."""
    if world > 1:
        import torch.distributed as dist
        flat = torch.empty(world * rec.flat.numel(), dtype=torch.float32, device=rec.flat.device)
        dist.all_gather_into_tensor(flat, rec.flat, group=group)      # rank r's record at [r*L, (r+1)*L)
        out = flat.view(world, rec.flat.numel())
    else:
        out = rec.flat[None]
    kp = out[:, :rec.nk].reshape(world * rec.slots * rec.max_kp, 2)                      # per-rank sections -> one array
    de = out[:, rec.nk:rec.nk + rec.nd].reshape(world * rec.slots * rec.max_kp, 128)
    cnt = out[:, rec.nk + rec.nd:].reshape(world * rec.slots).round().to(torch.int32)
    return kp.contiguous(), de.contiguous(), cnt


def global_frame_table(n_keyframes: int, world: int):
    """frame f -> (owner rank, local slot) under frames_of_rank's layout."""
    return {f: (f % world, f // world) for f in range(n_keyframes)}
