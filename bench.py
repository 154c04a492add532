#!/usr/bin/env python
"""bench.py - ALIKED+LightGlue frame-pairs/sec on synthetic KITTI-shaped frames (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|fp32x3|bf16] [--batch P]

Unit of work (BASELINE config 2): one frame-pair of the steady-state stream = 1 ALIKED extraction (frame t) + 1 LightGlue
match (t-1, t), 1241x376, 2048 kp.  A "step" is P = --batch consecutive pairs: P frames go through ONE batched extraction
(b2s_aliked_extract_batch, concurrent lanes) and the P pairs through ONE batched launch sequence of the matcher
(b2s_lightglue_match_batch_ex: the pair is a grid dimension of every kernel).

  value : device-resident throughput - u8 frames already in HBM, keypoints / descriptors / counts stay on the device
          between extraction and matching, two CUDA streams (extraction of step s+1 overlaps the matching of step s), no
          host synchronisation inside the timed region; CUDA events around the K steps, L2 flushed (256 MiB write, inside
          the timed region) before every step.
  e2e   : the same metric end to end with HOST buffers through the drop-in module: features_utils.FramePairStream.run
          (the sequence form of feature_extractor + feature_matcher for a recorded stream; identical results): host u8
          frames -> pinned staging -> H2D -> batched extraction + batched matching -> D2H of keypoints / descriptors /
          matches through pinned buffers -> lists of cv2.KeyPoint / cv2.DMatch, chunk c's host work overlapping chunk
          c + 1's GPU work.  Wall clock over the whole sequence.  e2e.per_call: the same frames one pair per call
          (feature_extractor + feature_matcher exactly as the reference's tracking loop issues them).
  roofline / roofline_gemm : the dominant kernel (attention) and the layer GEMMs timed live with CUDA events inside the
          library (b2s_lg_profile) in a separate pass right after the timed steps (the event pairs would perturb them);
          executed work is counted on the device.
  cpu_baseline : the CPU oracle (port of the reference's lightglue path) through the same API on the host cores, bounded
          sample; parity_check compares its cv2.DMatch sets with the GPU e2e leg's on the same frames.
  config3_window : BASELINE config 3 (16 keyframes, all 120 pairs) sharded over the N ranks with its one collective (a
          packed all_gather_into_tensor of the keyframe records), strong scaling, results checked against a local run.
Multi-GPU (torchrun): every rank streams its own contiguous chunk of frames (halo frame re-extracted locally, no
data-path collective) -> weak scaling; time = max over ranks.  --impl reference runs only the CPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NKP = 376, 1241, 2048
METRIC = "ALIKED+LightGlue frame-pairs/sec @1241x376, 2048 kp"
ADAPT_WEIGHTS = dict(token_bias=1.4, token_gain=6.0, match_bias=-5.0, match_gain=4.0)   # make early exit / pruning fire


def workload_config(src_a, src_l):
    """What BOTH arms run (identical dict in the two JSON lines); arm-specific details live under `run`."""
    return {"workload": "kitti_stream_1241x376_2048kp (BASELINE config 2)", "frame": f"{W}x{H} u8 BGR, synthetic (b200slam.synth)",
            "max_keypoints": NKP, "unit_of_work": "1 ALIKED-n16 extract + 1 LightGlue match (9 layers, adaptive depth 0.95 / width 0.99 on)",
            "api": "features_utils.feature_extractor + feature_matcher semantics (min_conf 0.7)", "weights": f"{src_a} / {src_l}"}


def lg_flops(m, n, L=9):
    """F_LG with the cross similarity counted once per layer."""
    return (2 * (m + n) * 128 * 256 + L * ((m + n) * 2_490_368 + 1024 * (m * m + n * n) + 1536 * m * n)
            + 2 * (m + n) * 256 ** 2 + 2 * m * n * 256)


def lg_gemm_flops(m, n, L=9):
    """The linear-layer share of F_LG (everything but the attention cores)."""
    return 2 * (m + n) * 128 * 256 + L * (m + n) * 2_490_368 + 2 * (m + n) * 256 ** 2 + 2 * m * n * 256


def aliked_flops(px=320 * 1024, n=NKP, M=16):
    return 20018.75 * px + n * (2 * 128 * 9 * 2 * M + 2 * (2 * M) ** 2 + 4 * M * 128 ** 2)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def _pos_pairs(kp0, kp1, matches):
    """cv2.DMatch list -> set of ((x0,y0),(x1,y1)) on a 1/8 px grid (keypoint ORDER may differ between two correct
    implementations when scores tie; positions do not)."""
    q = lambda kp: tuple(np.rint(np.array(kp.pt) * 8).astype(int).tolist())   # noqa: E731
    return {(q(kp0[m.queryIdx]), q(kp1[m.trainIdx])) for m in matches}


def cpu_reference_arm(steps, warmup, n_pairs_per_step=1, keep=None):
    """The reference's CPU path (oracle port; `lightglue` itself is not installable here) through
    feature_extractor/feature_matcher, all host threads.  Returns (pairs/s, ms_per_step, meta).  keep (dict, optional)
    receives {t: position-pair set of the matches of pair (t-1, t)} for the parity check."""
    import oracle  # noqa: F401  (test infrastructure; allowed here as the cpu baseline only)
    from oracle import features_utils as ofu
    from b200slam import weights, synth
    torch.set_num_threads(os.cpu_count() or 1)
    args = SimpleNamespace(use_lightglue=True, max_features=NKP, min_conf=0.7)
    sa, src_a = weights.load_aliked_state(allow_synthetic=True)
    sl, src_l = weights.load_lightglue_state(allow_synthetic=True)
    det, mat = ofu.init_feature_pipeline(args, sa, sl)
    frames = [synth.frame(t, H, W) for t in range((warmup + steps) * n_pairs_per_step + 1)]
    prev = ofu.feature_extractor(args, frames[0], det)
    times, nm, t_idx = [], [], 1
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for _ in range(n_pairs_per_step):
            cur = ofu.feature_extractor(args, frames[t_idx], det)
            ms = ofu.feature_matcher(args, prev[0], cur[0], prev[1], cur[1], mat)
            if keep is not None:
                keep[t_idx] = _pos_pairs(prev[0], cur[0], ms)
            prev = cur
            t_idx += 1
            nm.append(len(ms))
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    ms_per_step = 1e3 * float(np.mean(times))
    meta = {"kind": "port", "cores": torch.get_num_threads(),
            "sample": f"{steps * n_pairs_per_step} frame-pairs after {warmup * n_pairs_per_step} warm-up, CPU oracle via feature_extractor+feature_matcher",
            "mean_matches": float(np.mean(nm)), "torch": torch.__version__}
    return n_pairs_per_step * 1e3 / ms_per_step, ms_per_step, meta, (src_a, src_l)


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    v, ms, meta, (src_a, src_l) = cpu_reference_arm(args.steps, args.warmup, 1)
    meta["value"] = v
    meta["unit"] = "frame-pairs/s"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "frame-pairs/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": workload_config(src_a, src_l),
                      "run": {"pairs_per_step": 1, "device": "host CPU, all threads", "arithmetic": "torch fp32 (oneDNN/MKL)"},
                      "cpu_baseline": meta,
                      "e2e": {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


class Stream:
    """Device-resident frame stream: a ring of 3P feature slots; step s = frames sP+1 .. sP+P (slot (t-1) mod 3P), extracted
    in one batched call on stream A and matched (pairs (t-1, t)) in one batched launch sequence on stream B."""

    def __init__(self, det, frames_dev, P, dev, lanes):
        self.det, self.frames, self.P, self.dev, self.lanes = det, frames_dev, P, dev, lanes
        self.NS = 3 * P
        self.kp = torch.zeros((self.NS, NKP, 2), device=dev); self.de = torch.zeros((self.NS, NKP, 128), device=dev)
        self.sc = torch.zeros((self.NS, NKP), device=dev); self.cn = torch.zeros((self.NS,), dtype=torch.int32, device=dev)
        self.offs = (np.arange(self.NS, dtype=np.int64) * NKP).astype(np.int32)
        self.caps = np.full(self.NS, NKP, np.int32)
        prio = -1 if os.environ.get("B2S_BENCH_PRIO", "1") == "1" else 0   # the matcher's stream gets the higher priority
        self.s_ext, self.s_mat = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=prio)
        if os.environ.get("B2S_BENCH_SERIAL", "0") == "1":     # experiment: extraction and matching back to back on ONE stream
            self.s_ext = self.s_mat
        self.ext_done = [torch.cuda.Event() for _ in range(3)]
        self.mat_done = [torch.cuda.Event() for _ in range(3)]
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
        self.outs = [None, None, None]
        from b200slam import _lib
        self.fmt = _lib.IMG_BGR_U8_HWC

    def slot(self, t):
        return (t - 1) % self.NS

    def extract_step(self, s, flush=False):
        """frames sP+1 .. sP+P -> slots block (s mod 3); s = -1 extracts only frame 0 (the stream's first halo frame)."""
        with torch.cuda.stream(self.s_ext):
            if s >= 2:
                # block s mod 3 is being overwritten: its last slot is the halo frame of step s - 2's ... no: of the match of
                # step s - 2 (pair (t-1, t) with t the first frame of step s - 2 + 1), the most recent reader of this block
                self.s_ext.wait_event(self.mat_done[(s - 2) % 3])
            if flush:
                self.flush.fill_(s & 0xFF)                              # L2 flush before every timed step (inside the timed region)
            ts = [0] if s < 0 else list(range(s * self.P + 1, s * self.P + self.P + 1))
            b0 = self.slot(ts[0])
            imgs = [self.frames[t % len(self.frames)] for t in ts]
            n = len(ts)
            self.det.extract_batch_device(imgs, self.fmt, H, W, 3 * W, lanes=self.lanes,
                                          out=(self.kp[b0:b0 + n], self.de[b0:b0 + n], self.sc[b0:b0 + n], self.cn[b0:b0 + n]))
            if s >= 0:
                self.ext_done[s % 3].record(self.s_ext)

    def match_step(self, s, matcher):
        ts = range(s * self.P + 1, s * self.P + self.P + 1)
        pi = np.asarray([self.slot(t - 1) for t in ts], np.int32); pj = np.asarray([self.slot(t) for t in ts], np.int32)
        with torch.cuda.stream(self.s_mat):
            self.s_mat.wait_event(self.ext_done[s % 3])
            self.outs[s % 3] = matcher.match_batch_packed(self.kp.view(-1, 2), self.de.view(-1, 128), self.offs, pi, pj, stride=NKP,
                                                          out=self.outs[s % 3], counts=self.caps, counts_dev=self.cn)
            self.mat_done[s % 3].record(self.s_mat)
        return self.outs[s % 3]

    def run(self, matcher, steps, warm, world=1, sampler=None):
        """(warm + steps) steps; returns (total ms of the timed steps, launches inside them, match counts of the last step, wall s)."""
        cur = torch.cuda.current_stream(self.dev)
        self.s_ext.wait_stream(cur); self.s_mat.wait_stream(cur)
        self.extract_step(-1)
        self.extract_step(0)
        for s in range(warm):
            self.extract_step(s + 1)
            self.match_step(s, matcher)
        cur.wait_stream(self.s_ext); cur.wait_stream(self.s_mat)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        l0 = self.det.launches + matcher.launches
        if sampler is not None:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        e0.record(cur)
        self.s_ext.wait_event(e0); self.s_mat.wait_event(e0)
        # step `warm`'s frames were extracted before the timed region (pipeline prologue); symmetrically the timed region
        # extracts the frames of step warm + steps, which the NEXT region would match: K extractions + K matches are timed
        res = None
        for s in range(warm, warm + steps):
            self.extract_step(s + 1, flush=True)
            res = self.match_step(s, matcher)
        cur.wait_stream(self.s_ext); cur.wait_stream(self.s_mat)
        e1.record(cur)
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        launches = self.det.launches + matcher.launches - l0
        # precision fp32 = fp16x2 operand planes: a batch that left the fp16 range would report n = LG_RANGE (-2) and need
        # the bf16x3 re-run - that must not happen silently inside a timed region
        bad = sum(int((o["n"] < 0).sum()) for o in self.outs if o is not None)
        if bad:
            raise RuntimeError(f"{bad} pairs of the timed region left the fp16 operand range (LG_RANGE): re-run with --precision fp32x3")
        return e0.elapsed_time(e1), launches, res["n"].cpu().tolist(), res["stop"].cpu().tolist(), wall


def config3_window(det, mat, dev, rank, world, iters=5, warm=2):
    """BASELINE config 3: 16 keyframes (frames 0, 8, ..., 120), all 120 pairs, sharded over the ranks with ONE packed
    all_gather; strong scaling.  Every rank checks its pairs against a purely local run (all 16 frames extracted here)."""
    from b200slam import synth, window
    import torch.distributed as dist
    n_kf = 16
    frames = [torch.from_numpy(synth.frame(8 * t, H, W)).to(dev) for t in range(n_kf)]
    mat.reserve(NKP, mat.max_batch)
    res = None
    for _ in range(warm):
        res = window.match_keyframe_window(frames, det, mat, H, W, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gat = []
    e0.record()
    for _ in range(iters):
        tm = {}
        res = window.match_keyframe_window(frames, det, mat, H, W, rank, world, timing=tm)
        gat.append(tm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gather_us = 1e3 * float(np.mean([t["gather_start"].elapsed_time(t["gather_end"]) for t in gat]))
    # equality against a local (single-GPU semantics) run of this rank's pairs
    mine = {k: (v[0][: int(v[2])].cpu().numpy().copy(), v[1][: int(v[2])].cpu().numpy().copy()) for k, v in res.items()}
    ok = True
    if world > 1:
        loc = window.match_keyframe_window(frames, det, mat, H, W, 0, 1)
        torch.cuda.synchronize()
        for k, (m, s) in mine.items():
            lm, ls, ln = loc[k]
            ok &= bool(np.array_equal(m, lm[: int(ln)].cpu().numpy()) and np.array_equal(s, ls[: int(ln)].cpu().numpy()))
    nmatch = float(np.mean([len(v[0]) for v in mine.values()])) if mine else 0.0
    t = torch.tensor([ms, gather_us, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"workload": "keyframe_window_16kf_120pairs_2048kp (BASELINE config 3)", "scaling": "strong", "n_gpus": world,
            "pairs_per_s": 120.0 / (float(t[0]) / 1e3), "ms_per_window": float(t[0]), "all_gather_us": float(t[1]),
            "collective": "one all_gather_into_tensor of the flat (kpts | desc | count) record buffer, %.1f MB per rank" % (
                -(-n_kf // world) * NKP * 130 * 4 / 1e6),
            "equals_local_run": bool(float(t[2]) == 0.0), "mean_matches_per_pair": nmatch, "pairs_this_rank": len(mine),
            "timing": f"CUDA events over {iters} windows after {warm} warm-up, max over ranks; includes the 16 extractions"}


def run_ours(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - b200slam has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ncpu = os.cpu_count() or 1
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU on a shared host: keep each rank on its own slice of the cores (the e2e leg is host bound)
        per = max(1, ncpu // world)
        try:
            os.sched_setaffinity(0, set(range(local_rank * per, min(ncpu, (local_rank + 1) * per))))
        except Exception:
            pass
        torch.set_num_threads(per)
    from b200slam import features_utils as fu, synth, frontend, weights

    ns = SimpleNamespace(use_lightglue=True, max_features=NKP, min_conf=0.7, lg_precision=args.precision)
    sa, src_a = weights.load_aliked_state(allow_synthetic=True)
    sl, src_l = weights.load_lightglue_state(allow_synthetic=True)
    det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
    mat = frontend.LightGlue(weights=sl, device=dev, precision=args.precision, max_kp=NKP)

    K, Wm, P = args.steps, args.warmup, args.batch
    mat.reserve(NKP, P)
    n_frames = (K + Wm + 1) * P + 1
    t_base = rank * n_frames            # each rank owns its own contiguous chunk of the stream
    pool = 64                            # distinct frames kept in HBM (re-used cyclically)
    frames_np = [synth.frame(t_base + t, H, W) for t in range(min(n_frames, pool))]
    frames_dev = [torch.from_numpy(f).to(dev) for f in frames_np]
    stream = Stream(det, frames_dev, P, dev, args.lanes)

    sampler = ClockSampler(local_rank)
    total_ms, launches, mcounts, stops, wall = stream.run(mat, K, Wm, world, sampler)
    clocks = sampler.stop()
    mean_matches = float(np.mean(mcounts))

    # ---- per-kernel timing pass (outside the timed region: the event pairs around every attention / GEMM launch
    #      would perturb it): single stream, one batch of P pairs, CUDA events on the launching stream inside the library
    def kernel_pass(matcher, reps=3):
        stream.extract_step(-1); stream.extract_step(0)
        torch.cuda.synchronize()
        stream.match_step(0, matcher)
        torch.cuda.synchronize()
        matcher.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream.s_mat):
            e0.record()
            for _ in range(reps):
                r = stream.match_step(0, matcher)
            e1.record()
        torch.cuda.synchronize()
        attn_ms, attn_n = matcher.profile_read(0)
        gemm_ms, gemm_n = matcher.profile_read(1)
        self_w, cross_w = matcher.profile_work()
        matcher.profile(False)
        cn = stream.cn.cpu().numpy()
        ts = range(1, P + 1)
        sizes = [(int(cn[stream.slot(t - 1)]), int(cn[stream.slot(t)])) for t in ts]
        gemm_alg = reps * sum(lg_gemm_flops(m, n, int(L)) for (m, n), L in zip(sizes, r["stop"].cpu().tolist()))
        return dict(attn_ms=attn_ms, attn_n=attn_n, gemm_ms=gemm_ms, gemm_n=gemm_n, self_w=self_w, cross_w=cross_w,
                    match_ms=e0.elapsed_time(e1) / (reps * P), pairs=reps * P, gemm_alg=gemm_alg)

    kp_ = kernel_pass(mat)

    # ---- e2e through the drop-in API with host buffers, one pair per call like the reference's callers ----------
    e2e_pairs = max(4, min(K * P, 24))
    e2e_sets = {}
    ransac_leg = []
    prev = fu.feature_extractor(ns, frames_np[0], det)
    for t in range(1, 4):   # warm-up; rank 0 keeps these pairs' match sets for the parity check (frames 0..3 = the CPU leg's)
        cur = fu.feature_extractor(ns, frames_np[t % len(frames_np)], det)
        ms = fu.feature_matcher(ns, prev[0], cur[0], prev[1], cur[1], mat)
        e2e_sets[t] = _pos_pairs(prev[0], cur[0], ms)
        if rank == 0:
            ransac_leg.append((prev[0], cur[0], ms))
        prev = cur
    # ---- the stage right behind the matcher: filter_matches_ransac, reference body (cv2) vs OpenCV's RANSAC
    #      reproduced on the GPU (b2s_fm_cv_ransac_host) on the matches just produced ----------
    ransac = None
    if rank == 0 and ransac_leg:
        fu.filter_matches_ransac_gpu_cv2(*ransac_leg[0], 1.0)       # handle creation
        t_cv = t_gpu = 0.0
        same = True
        sizes = []
        for kpa, kpb, mm in ransac_leg:
            ta = time.perf_counter()
            keep_cv = fu.filter_matches_ransac_cv2(kpa, kpb, mm, 1.0)
            tb = time.perf_counter()
            keep_gpu = fu.filter_matches_ransac_gpu_cv2(kpa, kpb, mm, 1.0)
            tc = time.perf_counter()
            t_cv += tb - ta; t_gpu += tc - tb
            same &= [(m.queryIdx, m.trainIdx) for m in keep_cv] == [(m.queryIdx, m.trainIdx) for m in keep_gpu]
            sizes.append((len(mm), len(keep_cv)))
        ransac = {"pairs": len(ransac_leg), "matches_in_out": sizes, "identical_to_cv2": bool(same),
                  "cv2_ms_per_call": 1e3 * t_cv / len(ransac_leg), "gpu_cv2_ms_per_call": 1e3 * t_gpu / len(ransac_leg),
                  "what": "features_utils.filter_matches_ransac(kp1, kp2, matches, 1.0) on the per-call leg's first pairs: the reference's "
                          "cv2.findFundamentalMat call vs b2s_fm_cv_ransac_host (host lists in, surviving cv2.DMatch list out, "
                          "wall clock incl. the Python point gathering)"}
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for t in range(4, 4 + e2e_pairs):
        cur = fu.feature_extractor(ns, frames_np[t % len(frames_np)], det)
        ms = fu.feature_matcher(ns, prev[0], cur[0], prev[1], cur[1], mat)
        h2d += H * W * 3 + fu.last_match_h2d_bytes()
        d2h += len(cur[0]) * (2 + 128 + 1) * 4 + 4 + min(len(prev[0]), len(cur[0])) * 12 + 8
        prev = cur
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_matches = len(ms)

    # ---- e2e through the SEQUENCE form of the same API (features_utils.FramePairStream): host frames in, cv2 lists out,
    #      chunks of P frames, H2D / D2H through pinned buffers on a copy stream inside the timed region ----------
    seq_pairs = max(2 * P, min(4 * K * P, 256))     # a recorded sequence is long: pipeline fill / drain (one chunk each) must not dominate
    seq_frames = [frames_np[t % len(frames_np)] for t in range(seq_pairs + 1)]
    fps = fu.FramePairStream(ns, det, mat, batch=P, lanes=args.lanes)
    for _ in fps.run(seq_frames[: 2 * P + 1]):     # warm-up: allocations, graph capture of the lanes with the fused re-normalisation
        pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    fps.h2d_bytes = fps.d2h_bytes = 0
    t0 = time.perf_counter()
    seq_matches = 0
    for kps_t, des_t, m_t in fps.run(seq_frames):
        if m_t is not None:
            seq_matches = len(m_t)
    torch.cuda.synchronize()
    seq_s = time.perf_counter() - t0
    seq_h2d, seq_d2h = fps.h2d_bytes, fps.d2h_bytes
    if fps.range_fallback_pairs:
        raise RuntimeError("pairs of the timed sequence left the fp16 operand range")

    # ---- secondary legs on the same stream (reported next to the headline; this rank only) ----
    other = adaptive = None
    if not args.no_secondary:
        if args.precision == "fp32":
            mat16 = frontend.LightGlue(weights=sl, device=dev, precision="bf16", max_kp=NKP)
            mat16.reserve(NKP, P)
            ms16, _, mc16, _, _ = stream.run(mat16, max(2, K // 2), 2)
            other = {"precision": "bf16", "value": world * max(2, K // 2) * P / (ms16 / 1e3), "unit": "frame-pairs/s",
                     "ms_per_step": ms16 / max(2, K // 2), "mean_matches_per_pair": float(np.mean(mc16)),
                     "note": "operands rounded once to bf16 (>= 99 % match-set agreement, tests/test_gpu_tensorcore.py); every rank runs it, rank 0's time"}
            del mat16
        # weights that make the adaptive depth / width machinery fire (the seeded default weights run all 9 layers unpruned)
        sl_ad = weights.synthetic_lightglue_state(seed=0, **ADAPT_WEIGHTS)
        mat_ad = frontend.LightGlue(weights=sl_ad, device=dev, precision=args.precision, max_kp=NKP, filter_threshold=1e-6)
        mat_ad.reserve(NKP, P)
        msad, _, mcad, stad, _ = stream.run(mat_ad, max(2, K // 2), 2)
        adaptive = {"value": world * max(2, K // 2) * P / (msad / 1e3), "unit": "frame-pairs/s", "mean_stop_layer": float(np.mean(stad)),
                    "stop_layers_last_step": stad, "mean_matches_per_pair": float(np.mean(mcad)), "precision": args.precision,
                    "weights": f"synthetic(seed=0, {ADAPT_WEIGHTS}), filter_threshold 1e-6",
                    "note": "early exit and point pruning decided per pair on the device inside the batched launch sequence"}
        del mat_ad

    win = config3_window(det, mat, dev, rank, world) if not args.no_window else None

    t_total = torch.tensor([total_ms, e2e_s, seq_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max, seq_s_max = float(t_total[0]), float(t_total[1]), float(t_total[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * K * P / (total_ms_max / 1e3)
    e2e_value = world * e2e_pairs / e2e_s_max
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s"
    # dominant kernel: attention.  Work counted on the device (live sizes after pruning / early exit):
    #   self_w = sum nq*nk over self problems, cross_w = same over cross problems (both directions).
    # SURVEY 8d algorithmic FLOPs: self 1024 (M^2 + N^2) = 1024 self_w ; cross 1536 M N = 768 cross_w (similarity once);
    # executed: 1024 (self_w + cross_w) (the kernel forms Q K^T for both directions); the fp32 path issues 6 bf16
    # MMAs per product (three operand planes), the bf16 path 1.
    attn_s = kp_["attn_ms"] * 1e-3
    gemm_s = kp_["gemm_ms"] * 1e-3
    alg = 1024.0 * kp_["self_w"] + 768.0 * kp_["cross_w"]
    executed = 1024.0 * (kp_["self_w"] + kp_["cross_w"])
    issue_mult = {"fp32": 3, "fp32x3": 6, "bf16": 1}[args.precision]
    achieved = alg / attn_s / 1e12 if attn_s > 0 else None
    gemm_ach = kp_["gemm_alg"] / gemm_s / 1e12 if gemm_s > 0 else None
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json"))).get(args.precision, {})
    except Exception:
        pass
    out = {
        "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision != "bf16" else "bf16", "data": "synthetic",
        "config": workload_config(src_a, src_l),
        "run": {"pairs_per_step": P, "batch": f"{P} frames per batched extraction ({args.lanes or 4} concurrent CUDA-graph lanes), {P} pairs per batched matcher launch sequence",
                "l2": "flushed before every step (256 MiB write on the extraction stream, inside the timed region)",
                "pipeline": "2 CUDA streams: extraction of step s+1 overlaps the matching of step s; keypoint counts stay on the device (no host sync in the timed region)",
                "mean_matches_per_pair": mean_matches, "stop_layers_last_step": stops, "precision": args.precision,
                "arithmetic": {"fp32": "fp32-faithful on tcgen05: operands as two fp16 planes (x = h0 + 2^-11 h1), three cross products, fp32 accumulate; "
                                       "device-side fp16 range check, flagged batches re-run on bf16x3 planes",
                               "fp32x3": "fp32-faithful on tcgen05: operands as three bf16 planes, six cross products, fp32 accumulate",
                               "bf16": "bf16 operands on tcgen05, fp32 accumulate"}[args.precision],
                "launches_per_pair": launches / (K * P)},
        "gpu_launches": int(launches),
        "e2e": {"value": world * seq_pairs / seq_s_max, "unit": "frame-pairs/s", "h2d_bytes_per_step": int(seq_h2d / seq_pairs * P),
                "d2h_bytes_per_step": int(seq_d2h / seq_pairs * P), "pairs_timed": seq_pairs,
                "api": "features_utils.FramePairStream.run(host frames): the sequence form of feature_extractor + feature_matcher for a recorded "
                       "stream (BASELINE config 2 is one) - host u8 frames in, per frame (list[cv2.KeyPoint], np.float32 descriptors, list[cv2.DMatch]) "
                       f"out, identical to the per-call results (tests/test_gpu_e2e.py); chunks of {P} frames, H2D / D2H through pinned buffers on a copy "
                       "stream and the cv2 object construction all inside the timed wall-clock region",
                "matches_last_pair": seq_matches,
                "per_call": {"value": e2e_value, "unit": "frame-pairs/s", "h2d_bytes_per_step": int(h2d / e2e_pairs * P),
                             "d2h_bytes_per_step": int(d2h / e2e_pairs * P), "pairs_timed": e2e_pairs,
                             "api": "features_utils.feature_extractor + feature_matcher, one pair per call as the reference's tracking loop "
                                    "issues them (host numpy in, cv2 lists out; every call waits for its own copies and kernels)",
                             "matches_last_pair": e2e_matches}},
        "roofline": {"bound": "tensor",
                     "kernel": {"fp32": "k_attn_tc3<2> (fp32 on fp16x2 planes, tcgen05)", "fp32x3": "k_attn_tc3<3> (fp32 on bf16x3 planes, tcgen05)",
                                "bf16": "k_attn_tc (bf16, tcgen05)"}[args.precision],
                     "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                     "frac": (achieved / tensor_peak) if achieved else None,
                     "traffic": ncu.get("dram_bytes_per_launch"),
                     "peak_source": peak_src, "launches_timed": int(kp_["attn_n"]), "pairs_per_launch": P,
                     "avg_launch_ms": kp_["attn_ms"] / max(kp_["attn_n"], 1),
                     "algorithmic_flop_per_launch": alg / max(kp_["attn_n"], 1),
                     "executed_tflops": executed / attn_s / 1e12 if attn_s > 0 else None,
                     "mma_issued_tflops": issue_mult * executed / attn_s / 1e12 if attn_s > 0 else None,
                     "mma_issued_frac_of_peak": issue_mult * executed / attn_s / 1e12 / tensor_peak if attn_s > 0 else None,
                     "share_of_match": kp_["attn_ms"] / kp_["pairs"] / kp_["match_ms"],
                     "match_ms_per_pair_single_stream": kp_["match_ms"],
                     "timing": "CUDA events around every attention launch on the launching stream, separate pass of "
                               f"{kp_['pairs']} pairs after the timed region; work counted on the device (live sizes)",
                     "whole_pair_tflops": (lg_flops(NKP, NKP) + aliked_flops()) * K * P / (total_ms_max * 1e-3) / 1e12},
        "roofline_gemm": {"bound": "tensor", "kernel": "k_gemm_tcp (persistent tile scheduler, layer GEMMs + input / assignment projections)",
                          "achieved": gemm_ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": (gemm_ach / tensor_peak) if gemm_ach else None,
                          "traffic": ncu.get("gemm_dram_bytes_per_launch"), "peak_source": peak_src, "launches_timed": int(kp_["gemm_n"]),
                          "avg_launch_ms": kp_["gemm_ms"] / max(kp_["gemm_n"], 1),
                          "algorithmic_flop_per_launch": kp_["gemm_alg"] / max(kp_["gemm_n"], 1),
                          "mma_issued_frac_of_peak": issue_mult * gemm_ach / tensor_peak if gemm_ach else None,
                          "share_of_match": kp_["gemm_ms"] / kp_["pairs"] / kp_["match_ms"]},
        "clocks": clocks,
        "wall_s_timed_region": wall,
    }
    if other:
        out["other_precision"] = other
    if adaptive:
        out["adaptive_leg"] = adaptive
    if win:
        out["config3_window"] = win
    if ransac:
        out["ransac_filter"] = ransac
    if world == 1 and not args.no_cpu_baseline:
        keep = {}
        v, ms_cpu, meta, _ = cpu_reference_arm(steps=2, warmup=1, keep=keep)
        meta.update(value=v, unit="frame-pairs/s")
        out["cpu_baseline"] = meta
        agree = sum(len(keep[t] & e2e_sets[t]) for t in keep)
        tot = sum(len(keep[t] | e2e_sets[t]) for t in keep)
        out["parity_check"] = {"pairs": sorted(keep), "identical": all(keep[t] == e2e_sets[t] for t in keep), "agree": agree, "total": tot,
                               "what": "cv2.DMatch sets (by keypoint position) of the GPU drop-in e2e leg vs the CPU oracle leg on the same frames"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("B2S_PRECISION", "fp32"), choices=["fp32", "fp32x3", "bf16"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("B2S_BENCH_BATCH", "8")), help="pairs per step = per batched launch sequence")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("B2S_BENCH_LANES", "8")), help="concurrent extractor lanes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the bf16 / adaptive side measurements")
    ap.add_argument("--no-window", action="store_true", help="skip the config-3 keyframe-window sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
