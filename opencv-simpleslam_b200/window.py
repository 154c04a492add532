"""Keyframe-window all-pairs matching across GPUs (BASELINE config 3): every rank extracts the
keyframes f with f mod world == rank, ONE packed all_gather (NCCL over NVLink/NVSwitch; gloo in CPU tests) shares the
(keypoints, descriptors, count) records, then rank r matches the window pairs p with p mod world == r in batched launch
sequences (the pair is a grid dimension of every matcher kernel).  The reference runs this window through
`select_keyframe` / `triangulate_between_kfs_2view` one pair at a time on one device
(slam/core/keyframe_utils.py:153-154, triangulation_utils.py:131-132)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib, sharding


def match_keyframe_window(frames: List[torch.Tensor], det, mat, H: int, W: int, rank: int = 0, world: int = 1, group=None,
                          max_batch: int = 0, timing: dict | None = None, lanes: int = 8, check_range: bool = True
                          ) -> Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """frames: u8 BGR HWC CUDA tensors of ALL keyframes of the window (each rank only touches its own).
    Returns {(i, j): (matches int32 [stride,2], scores f32 [stride], n int32 [])} for the pairs this rank owns; tensors
    stay on the device (rows >= n are undefined).  timing (optional dict) receives CUDA events around the gather.
    check_range (precision='fp32' matcher only): read the match counts back once and re-run pairs that left the fp16
    operand range on the bf16x3 engine (LightGlue.resolve_range); False leaves n = LG_RANGE for the caller to handle."""
    dev = det.device
    n_kf, max_kp = len(frames), det.n_limit
    mine = sharding.frames_of_rank(n_kf, rank, world)
    slots = -(-n_kf // world)                               # same number of slots on every rank (padded with count 0)
    rec = sharding.WindowRecord(slots, max_kp, dev)
    if mine:
        # this rank's keyframes in ONE batched extraction (concurrent extractor lanes) straight into the record buffer
        cnt = torch.zeros((len(mine),), dtype=torch.int32, device=dev)
        det.extract_batch_device([frames[f] for f in mine], _lib.IMG_BGR_U8_HWC, H, W, 3 * W, lanes=lanes,
                                 out=(rec.kpts[:len(mine)], rec.desc[:len(mine)], None, cnt))
        rec.counts[:len(mine)].copy_(cnt)                      # int32 -> exact float32
    if timing is not None:
        timing["gather_start"] = torch.cuda.Event(enable_timing=True); timing["gather_end"] = torch.cuda.Event(enable_timing=True)
        timing["gather_start"].record()
    kp_all, de_all, counts = sharding.gather_window_records(rec, world, group)   # [world*slots*max_kp, 2|128], [world*slots]
    if timing is not None:
        timing["gather_end"].record()
    table = sharding.global_frame_table(n_kf, world)
    slot_of = lambda f: table[f][0] * slots + table[f][1]   # noqa: E731
    my_pairs = sharding.shard_pairs(sharding.window_pairs(n_kf), rank, world)
    if not my_pairs:
        return {}
    offs = (np.arange(world * slots, dtype=np.int64) * max_kp).astype(np.int32)
    pi = np.asarray([slot_of(i) for i, _ in my_pairs], np.int32)
    pj = np.asarray([slot_of(j) for _, j in my_pairs], np.int32)
    # the keypoint counts stay on the device (counts_dev): the host passes the slot capacity and never waits for them
    caps = np.full(world * slots, max_kp, np.int32)
    r = mat.match_batch_packed(kp_all, de_all, offs, pi, pj, stride=max_kp, max_batch=max_batch, counts=caps, counts_dev=counts)
    if check_range and mat.precision == "fp32":
        mat.resolve_range(r, kp_all, de_all, offs, pi, pj, max_batch, caps, counts)
    if check_range and bool((counts < 0).any().item()):      # a keyframe's fp16x2 convolutions left the fp16 range (any rank)
        raise _lib.B2SError(_lib.ALIKED_RANGE_MSG)
    return {pair: (r["matches"][p], r["scores"][p], r["n"][p]) for p, pair in enumerate(my_pairs)}
