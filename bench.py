#!/usr/bin/env python
"""bench.py - ALIKED+LightGlue frame-pairs/sec on synthetic KITTI-shaped frames (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

A "step" is PAIRS_PER_STEP frame-pairs of the steady-state stream (BASELINE config 2: extract
frame t, match (t-1, t)) = 1 ALIKED extraction + 1 LightGlue match per unit, 1241x376, 2048 kp.

  value : device-resident throughput - u8 frames already in HBM, keypoints/descriptors stay on the
          device between extraction and matching (b2s_aliked_extract / b2s_lightglue_match), timed
          with CUDA events per step on the launching stream, L2 flushed between steps.
  e2e   : the same metric through the drop-in API (features_utils.feature_extractor +
          feature_matcher) with HOST numpy buffers: H2D of the frame, D2H of keypoints/descriptors,
          H2D of both frames' features for the match, D2H of the matches; lists of
          cv2.KeyPoint/cv2.DMatch built. Wall clock bracketed by synchronisation.
  roofline : dominant kernel (attention) timed live with CUDA events inside the library
          (b2s_lg_profile) in a separate pass right after the timed steps (the event pairs would
          perturb them); executed work is counted on the device.
  cpu_baseline : the CPU oracle (port of the reference's lightglue path) through the same API on
          the host cores, bounded sample.  --impl reference runs only that arm.
Multi-GPU (torchrun): every rank streams its own contiguous chunk of frames (halo frame
re-extracted locally, no data-path collective) -> weak scaling; time = max over ranks.
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
PAIRS_PER_STEP = 8
METRIC = "ALIKED+LightGlue frame-pairs/sec @1241x376, 2048 kp"


def lg_flops(m, n, L=9):
    """F_LG with the cross similarity counted once per layer."""
    return (2 * (m + n) * 128 * 256 + L * ((m + n) * 2_490_368 + 1024 * (m * m + n * n) + 1536 * m * n)
            + 2 * (m + n) * 256 ** 2 + 2 * m * n * 256)


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


def cpu_reference_arm(steps, warmup, n_pairs_per_step=1):
    """The reference's CPU path (oracle port; `lightglue` itself is not installable here) through
    feature_extractor/feature_matcher, all host threads.  Returns (pairs/s, ms_per_step, meta)."""
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
            prev = cur
            t_idx += 1
            nm.append(len(ms))
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    ms_per_step = 1e3 * float(np.mean(times))
    meta = {"kind": "port", "cores": torch.get_num_threads(),
            "sample": f"{steps * n_pairs_per_step} frame-pairs after {warmup * n_pairs_per_step} warm-up, CPU oracle via feature_extractor+feature_matcher",
            "weights": f"{src_a} / {src_l}", "mean_matches": float(np.mean(nm)), "torch": torch.__version__}
    return n_pairs_per_step * 1e3 / ms_per_step, ms_per_step, meta


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    v, ms, meta = cpu_reference_arm(args.steps, args.warmup, 1)
    meta["value"] = v
    meta["unit"] = "frame-pairs/s"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "frame-pairs/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "kitti_stream_1241x376_2048kp (BASELINE config 2)", "pairs_per_step": 1},
                      "cpu_baseline": meta,
                      "e2e": {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_ours(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - b200slam has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from b200slam import features_utils as fu, synth, frontend, weights, _lib

    ns = SimpleNamespace(use_lightglue=True, max_features=NKP, min_conf=0.7, lg_precision=args.precision)
    sa, src_a = weights.load_aliked_state(allow_synthetic=True)
    sl, src_l = weights.load_lightglue_state(allow_synthetic=True)
    det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
    mat = frontend.LightGlue(weights=sl, device=dev, precision=args.precision, max_kp=NKP)

    K, Wm, P = args.steps, args.warmup, PAIRS_PER_STEP
    n_frames = (K + Wm) * P + 1
    t_base = rank * n_frames            # each rank owns its own contiguous chunk of the stream
    pool = 64                            # distinct frames kept in HBM (re-used cyclically)
    frames_np = [synth.frame(t_base + t, H, W) for t in range(min(n_frames, pool))]
    frames_dev = [torch.from_numpy(f).to(dev) for f in frames_np]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def device_run(matcher, steps, warm, sampler=None):
        """Streams (warm + steps) * P frame pairs through extract + match with everything resident in HBM.
        Software pipeline over two CUDA streams: while LightGlue matches (t-1, t) on stream B, ALIKED already
        extracts frame t+1 on stream A.  3 feature slots: t-1 and t are being matched while t+1 is written.
        Returns (per-step ms list, launches inside the timed steps, match counts, wall seconds)."""
        NS = 3
        kp_buf = [torch.empty((NKP, 2), device=dev) for _ in range(NS)]
        de_buf = [torch.empty((NKP, 128), device=dev) for _ in range(NS)]
        n_host = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(NS)]
        counts = [0] * NS
        # B2S_BENCH_PRIO (default 1): the matcher's stream gets the higher priority (its kernels leave 20 of the 148 SMs idle for the extractor)
        prio = -1 if os.environ.get("B2S_BENCH_PRIO", "1") == "1" else 0
        s_ext, s_mat = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=prio)
        ext_done = [torch.cuda.Event() for _ in range(NS)]
        mat_done = torch.cuda.Event()

        def enqueue_extract(t):
            slot = t % NS
            with torch.cuda.stream(s_ext):
                s_ext.wait_event(mat_done)     # slot t%3 was last read by match(t-3, t-2): already enqueued before
                kp, de, _, n = det.extract_device(frames_dev[t % len(frames_dev)], _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
                kp_buf[slot].copy_(kp, non_blocking=True)
                de_buf[slot].copy_(de, non_blocking=True)
                n_host[slot].copy_(n, non_blocking=True)
                ext_done[slot].record(s_ext)

        def finish_extract(t):
            slot = t % NS
            ext_done[slot].synchronize()       # the keypoint count sizes the matcher's launch
            counts[slot] = int(n_host[slot][0])

        def device_step(t0):
            res = None
            for i in range(P):
                t = t0 + i
                finish_extract(t)
                enqueue_extract(t + 1)
                cur, prv = t % NS, (t - 1) % NS
                with torch.cuda.stream(s_mat):
                    s_mat.wait_event(ext_done[cur])
                    res = matcher.match_device(kp_buf[prv][:counts[prv]], de_buf[prv][:counts[prv]],
                                               kp_buf[cur][:counts[cur]], de_buf[cur][:counts[cur]], full=False)
                    mat_done.record(s_mat)
            return res

        def join_streams():
            cur = torch.cuda.current_stream()
            cur.wait_stream(s_ext); cur.wait_stream(s_mat)

        enqueue_extract(0); finish_extract(0); enqueue_extract(1)
        for s in range(warm):
            device_step(1 + s * P)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = det.launches + matcher.launches
        if sampler is not None:
            sampler.start()
        evs, mcounts = [], []
        wall0 = time.perf_counter()
        for s in range(steps):
            join_streams()
            flush.fill_(s & 0xFF)                       # L2 flush between timed iterations (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s_ext.wait_event(e0); s_mat.wait_event(e0)
            res = device_step(1 + (warm + s) * P)
            join_streams()
            e1.record()
            evs.append((e0, e1))
            mcounts.append(res["n"])
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        launches = det.launches + matcher.launches - l0
        return [a.elapsed_time(b) for a, b in evs], launches, [int(c.item()) for c in mcounts], wall

    sampler = ClockSampler(local_rank)
    step_ms, launches, mcounts, wall = device_run(mat, K, Wm, sampler)
    clocks = sampler.stop()
    total_ms = float(sum(step_ms))
    mean_matches = float(np.mean(mcounts))

    # ---- per-kernel timing pass (outside the timed region: the event pairs around every attention / GEMM launch
    #      would perturb it): single stream, the same frames, CUDA events on the launching stream inside the library
    def kernel_pass(matcher, pairs=4):
        feats = []
        for t in range(pairs + 1):
            kp, de, _, n = det.extract_device(frames_dev[t % len(frames_dev)], _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
            torch.cuda.synchronize()
            k = int(n.item())
            feats.append((kp[:k].clone(), de[:k].clone()))
        for t in range(2):
            matcher.match_device(feats[t][0], feats[t][1], feats[t + 1][0], feats[t + 1][1], full=False)
        torch.cuda.synchronize()
        matcher.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(pairs):
            matcher.match_device(feats[t][0], feats[t][1], feats[t + 1][0], feats[t + 1][1], full=False)
        e1.record()
        torch.cuda.synchronize()
        attn_ms, attn_n = matcher.profile_read(0)
        gemm_ms, gemm_n = matcher.profile_read(1)
        self_w, cross_w = matcher.profile_work()
        matcher.profile(False)
        return dict(attn_ms=attn_ms, attn_n=attn_n, gemm_ms=gemm_ms, gemm_n=gemm_n, self_w=self_w, cross_w=cross_w,
                    match_ms=e0.elapsed_time(e1) / pairs, pairs=pairs)

    kp_ = kernel_pass(mat)

    # ---- e2e through the drop-in API with host buffers -------------------------------------
    e2e_pairs = max(4, min(K * P, 24))
    prev = fu.feature_extractor(ns, frames_np[0], det)
    for t in range(1, 3):   # warm-up
        cur = fu.feature_extractor(ns, frames_np[t % len(frames_np)], det)
        fu.feature_matcher(ns, prev[0], cur[0], prev[1], cur[1], mat)
        prev = cur
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for t in range(3, 3 + e2e_pairs):
        cur = fu.feature_extractor(ns, frames_np[t % len(frames_np)], det)
        ms = fu.feature_matcher(ns, prev[0], cur[0], prev[1], cur[1], mat)
        h2d += H * W * 3 + (len(prev[0]) + len(cur[0])) * (2 + 128) * 4
        d2h += len(cur[0]) * (2 + 128 + 1) * 4 + 4 + min(len(prev[0]), len(cur[0])) * 12 + 4
        prev = cur
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_matches = len(ms)

    # ---- secondary: the bf16 matcher on the same stream (reported next to the fp32 headline) ----
    other = None
    if args.precision == "fp32" and not args.no_secondary:
        mat16 = frontend.LightGlue(weights=sl, device=dev, precision="bf16", max_kp=NKP)
        ms16, _, mc16, _ = device_run(mat16, max(2, K // 2), 2)
        other = {"precision": "bf16", "value": world * len(ms16) * P / (float(sum(ms16)) / 1e3), "unit": "frame-pairs/s",
                 "ms_per_step": float(np.mean(ms16)), "mean_matches_per_pair": float(np.mean(mc16)),
                 "note": "operands rounded once to bf16 (>= 99 % match-set agreement, tests/test_gpu_tensorcore.py); this rank only"}
        del mat16

    t_total = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t_total[0]), float(t_total[1])
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
    alg = 1024.0 * kp_["self_w"] + 768.0 * kp_["cross_w"]
    executed = 1024.0 * (kp_["self_w"] + kp_["cross_w"])
    issue_mult = {"fp32": 6, "bf16": 1, "fp32_simt": 0}[args.precision]
    achieved = alg / attn_s / 1e12 if attn_s > 0 else None
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json"))).get(args.precision, {})
    except Exception:
        pass
    out = {
        "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision != "bf16" else "bf16", "data": "synthetic",
        "config": {"workload": "kitti_stream_1241x376_2048kp (BASELINE config 2)", "pairs_per_step": P,
                   "unit_of_work": "1 ALIKED-n16 extract + 1 LightGlue match (9 layers, adaptive depth/width on)",
                   "l2": "flushed between steps (256 MiB write)", "weights": f"{src_a} / {src_l}",
                   "pipeline": "2 CUDA streams: ALIKED extract(t+1) overlaps LightGlue match(t-1,t); the matcher stream has the higher priority",
                   "mean_matches_per_pair": mean_matches, "precision": args.precision,
                   "arithmetic": {"fp32": "fp32-faithful on tcgen05: operands as three bf16 planes, six cross products, fp32 accumulate",
                                  "bf16": "bf16 operands on tcgen05, fp32 accumulate", "fp32_simt": "fp32 FMA on CUDA cores"}[args.precision]},
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "h2d_bytes_per_step": int(h2d / e2e_pairs * P),
                "d2h_bytes_per_step": int(d2h / e2e_pairs * P), "pairs_timed": e2e_pairs,
                "api": "features_utils.feature_extractor + feature_matcher (host numpy in, cv2 lists out)",
                "matches_last_pair": e2e_matches},
        "roofline": {"bound": "tensor",
                     "kernel": {"fp32": "k_attn_tc3 (fp32 on bf16x3 planes, tcgen05)", "bf16": "k_attn_tc (bf16, tcgen05)",
                                "fp32_simt": "k_attn_fp32 (CUDA cores)"}[args.precision],
                     "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                     "frac": (achieved / tensor_peak) if achieved else None,
                     "traffic": ncu.get("dram_bytes_per_launch"),
                     "peak_source": peak_src, "launches_timed": int(kp_["attn_n"]),
                     "avg_launch_ms": kp_["attn_ms"] / max(kp_["attn_n"], 1),
                     "algorithmic_flop_per_launch": alg / max(kp_["attn_n"], 1),
                     "executed_tflops": executed / attn_s / 1e12 if attn_s > 0 else None,
                     "mma_issued_tflops": issue_mult * executed / attn_s / 1e12 if attn_s > 0 else None,
                     "mma_issued_frac_of_peak": issue_mult * executed / attn_s / 1e12 / tensor_peak if attn_s > 0 else None,
                     "share_of_match": kp_["attn_ms"] / kp_["pairs"] / kp_["match_ms"],
                     "gemm_share_of_match": kp_["gemm_ms"] / kp_["pairs"] / kp_["match_ms"],
                     "match_ms_single_stream": kp_["match_ms"],
                     "timing": "CUDA events around every attention launch on the launching stream, separate pass of "
                               f"{kp_['pairs']} matches after the timed region; work counted on the device (live sizes)",
                     "whole_pair_tflops": (lg_flops(NKP, NKP) + aliked_flops()) * K * P / (total_ms * 1e-3) / 1e12},
        "clocks": clocks,
        "wall_s_timed_region": wall,
    }
    if other:
        out["other_precision"] = other
    if world == 1 and not args.no_cpu_baseline:
        v, ms_cpu, meta = cpu_reference_arm(steps=3, warmup=1)
        meta.update(value=v, unit="frame-pairs/s")
        out["cpu_baseline"] = meta
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("B2S_PRECISION", "fp32"), choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the bf16 side measurement of the fp32 run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
