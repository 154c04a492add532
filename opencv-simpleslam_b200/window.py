"""Keyframe-window all-pairs matching across GPUs (BASELINE config 3): every rank extracts the
keyframes f with f mod world == rank, ONE all_gather (NCCL over NVLink/NVSwitch; gloo in CPU tests) shares the
(count, keypoints, descriptors) records, then rank r matches the window pairs p with p mod world == r.  The reference
runs this window through `select_keyframe` / `triangulate_between_kfs_2view` one pair at a time on one device
(slam/core/keyframe_utils.py:153-154, triangulation_utils.py:131-132)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import _lib, sharding


def match_keyframe_window(frames: List[torch.Tensor], det, mat, H: int, W: int, rank: int = 0, world: int = 1, group=None
                          ) -> Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]]:
    """frames: u8 BGR HWC CUDA tensors of ALL keyframes of the window (each rank only touches its own).
    Returns {(i, j): (matches int32 [K,2], scores f32 [K])} for the pairs this rank owns; tensors stay on the device."""
    dev = det.device
    n_kf, max_kp = len(frames), det.n_limit
    mine = sharding.frames_of_rank(n_kf, rank, world)
    slots = -(-n_kf // world)                               # same F_local on every rank (padded with count 0)
    kp = torch.zeros((slots, max_kp, 2), device=dev)
    de = torch.zeros((slots, max_kp, 128), device=dev)
    cnt = torch.zeros((slots,), dtype=torch.int32, device=dev)
    for s, f in enumerate(mine):
        k, d, _, n = det.extract_device(frames[f], _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
        kp[s].copy_(k); de[s].copy_(d); cnt[s:s + 1].copy_(n)
    if world > 1:
        gk, gd, gc = sharding.gather_window_features(kp, de, cnt, group)
    else:
        gk, gd, gc = [kp], [de], [cnt]
    counts = torch.stack(gc).cpu()                          # one small D2H: the counts size the matcher launches
    table = sharding.global_frame_table(n_kf, world)
    out = {}
    for (i, j) in sharding.shard_pairs(sharding.window_pairs(n_kf), rank, world):
        (ri, si), (rj, sj) = table[i], table[j]
        m, n = int(counts[ri, si]), int(counts[rj, sj])
        r = mat.match_device(gk[ri][si, :m], gd[ri][si, :m], gk[rj][sj, :n], gd[rj][sj, :n], full=False)
        out[(i, j)] = (r["matches"], r["scores"], r["n"])
    return out
