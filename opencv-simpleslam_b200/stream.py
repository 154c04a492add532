"""Sequence form of the drop-in frontend: consecutive-frame matching of a HOST frame stream, pipelined.

The reference's tracking loop (main_revamped.py:313-325) calls `feature_extractor(frame_t)` and
`feature_matcher(frame_{t-1}, frame_t)` one frame at a time; each call has to wait for the GPU and for its own
host <-> device copies.  When the frames are known ahead (a recorded sequence: KITTI / TUM / EuRoC loaders of the
reference, BASELINE configs 2 and 5) the same results can be produced in chunks of B frames:

    pinned staging -> H2D (copy stream) -> b2s_aliked_extract_batch (B frames on concurrent extractor lanes)
    -> b2s_lightglue_match_batch_ex (the B consecutive pairs, one launch sequence) -> D2H through pinned buffers
    (copy stream) -> cv2.KeyPoint / cv2.DMatch lists built on the host WHILE the GPU works on the next chunk.

`FramePairStream.run(frames)` yields, per frame, exactly what the two reference calls return:
    (keypoints: list[cv2.KeyPoint], descriptors: np.float32 [N,128], matches to the previous frame: list[cv2.DMatch])
(the first frame has no previous frame: its match list is None).  Results are identical to calling
`features_utils.feature_extractor` / `feature_matcher` frame by frame (tests/test_gpu_e2e.py): same kernels, same
descriptor re-normalisation, same min_conf filter, same ordering.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


class FramePairStream:
    def __init__(self, args, detector, matcher, batch: int = 8, lanes: int = 8):
        """args: the reference's CLI namespace (use_lightglue, min_conf, optional array_native); detector / matcher:
        the objects returned by `init_feature_pipeline` (b200slam.ALIKED / LightGlue on the same device)."""
        if not getattr(args, "use_lightglue", True):
            raise ValueError("FramePairStream serves the LightGlue branch only")
        if detector.device != matcher.device:
            raise ValueError("detector and matcher live on different devices")
        self.args, self.det, self.mat = args, detector, matcher
        self.B, self.lanes = int(batch), int(lanes)
        self.dev = detector.device
        self.nkp = detector.n_limit
        self.min_conf = float(getattr(args, "min_conf", 0.7))
        self.array_native = bool(getattr(args, "array_native", False))
        self._shape = None
        self.h2d_bytes = self.d2h_bytes = 0
        self.range_fallback_pairs = 0

    # ------------------------------------------------------------------------------------------------------------
    def _alloc(self, H: int, W: int):
        B, N, dev = self.B, self.nkp, self.dev
        S = 3 * B                                             # feature slots on the device (frame t lives in slot t % S)
        with torch.cuda.device(dev):
            self.s_main = torch.cuda.Stream(dev)
            self.s_io = torch.cuda.Stream(dev)
            self.pin_img = [torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
            self.dev_img = [torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.kp = torch.zeros((S, N, 2), dtype=torch.float32, device=dev)
            self.de = torch.zeros((S, N, 128), dtype=torch.float32, device=dev)
            self.cn = torch.zeros((S,), dtype=torch.int32, device=dev)
            self.offs = (np.arange(S, dtype=np.int64) * N).astype(np.int32)
            self.caps = np.full(S, N, np.int32)
            # per-chunk results on the device and their pinned mirrors: one packed int32 buffer per chunk
            #   [cn B | n B | stop B | kp B*N*2 | de B*N*128 | matches B*N*2 | scores B*N]
            self.n_words = 3 * B + B * N * (2 + 128 + 2 + 1)
            self.out_dev = [torch.zeros(self.n_words, dtype=torch.int32, device=dev) for _ in range(2)]
            self.out_pin = [torch.zeros(self.n_words, dtype=torch.int32).pin_memory() for _ in range(2)]
            self.ev_h2d = [torch.cuda.Event() for _ in range(2)]
            self.ev_done = [torch.cuda.Event() for _ in range(2)]
            self.ev_out = [torch.cuda.Event() for _ in range(2)]
            self.ev_img_free = [torch.cuda.Event() for _ in range(2)]
        self.mat.reserve(N, B)
        self._shape = (H, W)

    def _views(self, buf: torch.Tensor):
        B, N = self.B, self.nkp
        o = 3 * B
        f = buf.view(torch.float32)
        v = {"cn": buf[0:B], "n": buf[B:2 * B], "stop": buf[2 * B:3 * B]}
        v["kp"] = f[o:o + B * N * 2].view(B, N, 2); o += B * N * 2
        v["de"] = f[o:o + B * N * 128].view(B, N, 128); o += B * N * 128
        v["matches"] = buf[o:o + B * N * 2].view(B, N, 2); o += B * N * 2
        v["scores"] = f[o:o + B * N].view(B, N)
        return v

    # ------------------------------------------------------------------------------------------------------------
    def _enqueue(self, c: int, frames: List[np.ndarray], t0: int):
        """All GPU work of chunk c (frames t0 .. t0 + len(frames) - 1) - no host synchronisation."""
        B, N = self.B, self.nkp
        H, W = self._shape
        nb, k = len(frames), c % 2
        pin, dimg = self.pin_img[k], self.dev_img[k]
        if c >= 2:
            self.ev_h2d[k].synchronize()                       # the H2D of chunk c - 2 has read this pinned buffer (long ago)
        pn = pin.numpy()
        for i, f in enumerate(frames):
            if f.shape != (H, W, 3) or f.dtype != np.uint8:
                raise ValueError("FramePairStream: all frames must be u8 BGR of the same size")
            np.copyto(pn[i], f)
        S = 3 * B
        slots = [(t0 + i) % S for i in range(nb)]
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.s_io):
                if c >= 2:
                    self.s_io.wait_event(self.ev_img_free[k])  # extraction of chunk c - 2 has consumed dev_img[k]
                dimg[:nb].copy_(pin[:nb], non_blocking=True)
                self.ev_h2d[k].record(self.s_io)
            self.h2d_bytes += nb * H * W * 3
            with torch.cuda.stream(self.s_main):
                self.s_main.wait_event(self.ev_h2d[k])
                if c >= 2:
                    self.s_main.wait_event(self.ev_out[k])     # the D2H of chunk c - 2 has read out_dev[k]
                od = self._views(self.out_dev[k])
                # slots are contiguous runs (at most two: the ring wraps)
                runs = []
                i = 0
                while i < nb:
                    j = i
                    while j + 1 < nb and slots[j + 1] == slots[j] + 1:
                        j += 1
                    runs.append((i, slots[i], j - i + 1))
                    i = j + 1
                for i, s0, cnt in runs:
                    self.det.extract_batch_device([dimg[q] for q in range(i, i + cnt)], _lib.IMG_BGR_U8_HWC, H, W, 3 * W, lanes=self.lanes,
                                                  out=(self.kp[s0:s0 + cnt], self.de[s0:s0 + cnt], None, self.cn[s0:s0 + cnt]))
                self.ev_img_free[k].record(self.s_main)
                first = 1 if t0 == 0 else 0                    # frame 0 has no predecessor
                pi = np.asarray([(t0 + i - 1) % S for i in range(first, nb)], np.int32)
                pj = np.asarray([(t0 + i) % S for i in range(first, nb)], np.int32)
                od["n"].zero_()
                if len(pi):
                    mo = {"matches": od["matches"][first:nb], "scores": od["scores"][first:nb], "n": od["n"][first:nb], "stop": od["stop"][first:nb]}
                    self.mat.match_batch_packed(self.kp.view(-1, 2), self.de.view(-1, 128), self.offs, pi, pj, stride=N, out=mo,
                                                counts=self.caps, counts_dev=self.cn)
                for i, s0, cnt in runs:                        # plain device-to-device slices: nothing here may touch the host
                    od["cn"][i:i + cnt].copy_(self.cn[s0:s0 + cnt])
                    od["kp"][i:i + cnt].copy_(self.kp[s0:s0 + cnt])
                    od["de"][i:i + cnt].copy_(self.de[s0:s0 + cnt])
                self.ev_done[k].record(self.s_main)
            with torch.cuda.stream(self.s_io):
                self.s_io.wait_event(self.ev_done[k])
                self.out_pin[k].copy_(self.out_dev[k], non_blocking=True)
                self.ev_out[k].record(self.s_io)
            self.d2h_bytes += self.n_words * 4
        return nb, first

    def _collect(self, c: int, nb: int, first: int, t0: int, prev):
        """Host side of chunk c: wait for its D2H, build the reference's return types.  prev = (kp_arr, des) of frame t0 - 1."""
        from . import features_utils as fu
        from .containers import DMatchArray, KeyPointArray
        k = c % 2
        self.ev_out[k].synchronize()
        v = self._views(self.out_pin[k])
        cn = v["cn"].numpy(); nm = v["n"].numpy()
        kp_np = v["kp"].numpy(); de_np = v["de"].numpy(); mt_np = v["matches"].numpy(); sc_np = v["scores"].numpy()
        out = []
        for i in range(nb):
            n = int(cn[i])
            if n == _lib.ALIKED_RANGE:
                raise _lib.B2SError(_lib.ALIKED_RANGE_MSG)
            kp_arr = kp_np[i, :n].copy()
            des = de_np[i, :n].copy()
            kps = KeyPointArray(kp_arr) if self.array_native else fu._convert_lg_kps_to_opencv(kp_arr)
            matches = None
            if i >= first or t0 > 0:
                if prev is None or n == 0 or len(prev[1]) == 0:
                    matches = []
                else:
                    m = int(nm[i])
                    if m == _lib.LG_RANGE:                     # fp16x2 engine left its range: this pair goes through the bf16x3 engine
                        self.range_fallback_pairs += 1
                        raw = self.mat.fallback().match_mixed(prev[0], prev[1], kp_arr, des)
                        pairs, scores = raw["matches"], raw["scores"]
                    else:
                        pairs, scores = mt_np[i, :m], sc_np[i, :m]
                    keep = scores > np.float32(self.min_conf)
                    matches = DMatchArray(pairs[keep].copy()) if self.array_native else fu._convert_lg_matches_to_opencv(pairs[keep])
            out.append((kps, des, matches))
            prev = (kp_arr, des)
        return out, prev

    # ------------------------------------------------------------------------------------------------------------
    def run(self, frames: Iterable[np.ndarray]) -> Iterator[Tuple[list, np.ndarray, Optional[list]]]:
        """frames: iterable of u8 BGR HWC host arrays of one size.  Yields (keypoints, descriptors, matches_to_previous)."""
        it = iter(frames)
        pending = None        # (c, nb, first, t0) of the chunk whose GPU work is in flight
        prev = None
        c = t = 0
        check(lib.b2s_aliked_set_batch_renorm(self.det._handle, 1e-8), "b2s_aliked_set_batch_renorm")   # features_utils.py:100
        try:
            while True:
                chunk = []
                for f in it:
                    chunk.append(f)
                    if len(chunk) == self.B:
                        break
                if chunk:
                    if self._shape is None:
                        self._alloc(chunk[0].shape[0], chunk[0].shape[1])
                    nb, first = self._enqueue(c, chunk, t)
                    cur = (c, nb, first, t)
                    c += 1; t += nb
                else:
                    cur = None
                if pending is not None:
                    res, prev = self._collect(*pending, prev)
                    for r in res:
                        yield r
                pending = cur
                if cur is None:
                    break
        finally:
            check(lib.b2s_aliked_set_batch_renorm(self.det._handle, 0.0), "b2s_aliked_set_batch_renorm")
            if self._shape is not None:
                torch.cuda.synchronize(self.dev)


def match_sequence(args, frames: Iterable[np.ndarray], detector, matcher, batch: int = 8, lanes: int = 8):
    """One-shot form: list of (keypoints, descriptors, matches_to_previous_frame) for a host frame sequence."""
    return list(FramePairStream(args, detector, matcher, batch, lanes).run(frames))
