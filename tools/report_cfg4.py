"""BASELINE config 4: ALIKED-n32 at 1920x1080, 4096 keypoints + LightGlue at full depth (depth_confidence = -1,
width_confidence = -1): fp32 (bf16x3 planes) vs bf16 tolerance report, both precisions timed.  Prints one JSON line
(committed as profiles/r1_cfg4_report.json).  Run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib, frontend, synth, weights   # noqa: E402

H, W, NKP = 1080, 1920, 4096
dev = torch.device("cuda", 0)
sa = weights.synthetic_aliked_state("aliked-n32", seed=0)
sl = weights.synthetic_lightglue_state(seed=0)
det = frontend.ALIKED(model_name="aliked-n32", max_num_keypoints=NKP, weights=sa, device=dev)
frames = [torch.from_numpy(synth.frame(t, H, W)).to(dev) for t in (0, 1)]
feats = []
for f in frames:
    kp, de, sc, n = det.extract_device(f, _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
    torch.cuda.synchronize()
    k = int(n.item())
    feats.append((kp[:k].clone(), de[:k].clone()))


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ext_ms = timeit(lambda: det.extract_device(frames[1], _lib.IMG_BGR_U8_HWC, H, W, 3 * W))
res, ms = {}, {}
for prec in ("fp32", "bf16"):
    mat = frontend.LightGlue(weights=sl, device=dev, precision=prec, max_kp=NKP, depth_confidence=-1, width_confidence=-1)
    run = lambda: mat.match_device(feats[0][0], feats[0][1], feats[1][0], feats[1][1], full=True)   # noqa: E731
    r = run(); torch.cuda.synchronize()
    nm = int(r["n"].item())
    res[prec] = {"matches": r["matches"][:nm].cpu().numpy(), "scores": r["scores"][:nm].cpu().numpy(),
                 "matches0": r["matches0"].cpu().numpy(), "ms0": r["matching_scores0"].cpu().numpy(), "stop": int(r["stop"].item())}
    ms[prec] = timeit(run)
a, b = res["fp32"], res["bf16"]
sa_, sb_ = set(map(tuple, a["matches"].tolist())), set(map(tuple, b["matches"].tolist()))
both = sa_ & sb_
agree = float((a["matches0"] == b["matches0"]).mean())
ia = {tuple(m): s for m, s in zip(a["matches"].tolist(), a["scores"].tolist())}
ib = {tuple(m): s for m, s in zip(b["matches"].tolist(), b["scores"].tolist())}
rel = max(abs(ia[k] - ib[k]) / max(abs(ia[k]), 1e-9) for k in both) if both else None
F = 2 * (2 * NKP) * 128 * 256 + 9 * ((2 * NKP) * 2_490_368 + 1024 * 2 * NKP * NKP + 1536 * NKP * NKP) + 2 * (2 * NKP) * 256 ** 2 + 2 * NKP * NKP * 256
print(json.dumps({
    "workload": "BASELINE config 4: ALIKED-n32 1920x1080, 4096 kp, LightGlue full depth (no early exit, no pruning)",
    "keypoints": [len(feats[0][0]), len(feats[1][0])], "executed_layers": [a["stop"], b["stop"]],
    "aliked_n32_extract_ms": round(ext_ms, 3),
    "lightglue_ms": {k: round(v, 3) for k, v in ms.items()},
    "lightglue_algorithmic_gflop": round(F / 1e9, 1),
    "lightglue_tflops": {k: round(F / v / 1e9, 1) for k, v in ms.items()},
    "matches": {"fp32": len(sa_), "bf16": len(sb_), "common": len(both)},
    "match_set_jaccard": round(len(both) / max(len(sa_ | sb_), 1), 5),
    "matches0_agreement": round(agree, 5),
    "score_max_rel_err_on_common_matches": rel,
    "keypoint_sets": "identical by construction (ALIKED has one precision: fp32 on bf16x3 planes; the bf16 switch applies to the LightGlue layers)",
    "target": "match-set agreement >= 99 % (BASELINE.json north_star)"}))
