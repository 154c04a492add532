"""Host-side breakdown of the drop-in API, one pair per call (run under gpurun): python tools/profile_e2e.py [fp32|bf16]"""
import os, sys, time
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import features_utils as fu, synth, frontend, weights

H, W, NKP = 376, 1241, 2048
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
ns = SimpleNamespace(use_lightglue=True, max_features=NKP, min_conf=0.7, lg_precision=prec)
sa, _ = weights.load_aliked_state(allow_synthetic=True); sl, _ = weights.load_lightglue_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device="cuda:0")
mat = frontend.LightGlue(weights=sl, device="cuda:0", precision=prec, max_kp=NKP)
frames = [synth.frame(t, H, W) for t in range(14)]
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()
prev = fu.feature_extractor(ns, frames[0], det)
for rep in range(2):
    T.clear()
    for t in range(1, 11):
        t0 = time.perf_counter()
        seen = {}
        def on_kp(k):
            seen["t_kp"] = time.perf_counter(); seen["kp"] = k.copy()
            r = fu._convert_lg_kps_to_opencv(k); seen["t_list"] = time.perf_counter(); return r
        kp, des, _ = det.extract_host_split(frames[t], on_kp, desc_renorm_eps=1e-8)
        t1 = time.perf_counter()
        T["extract: begin -> keypoints on host"] = T.get("extract: begin -> keypoints on host", 0) + seen["t_kp"] - t0
        T["extract: KeyPoint list (overlaps SDDH)"] = T.get("extract: KeyPoint list (overlaps SDDH)", 0) + seen["t_list"] - seen["t_kp"]
        T["extract: finish (desc D2H)"] = T.get("extract: finish (desc D2H)", 0) + t1 - seen["t_list"]
        t0 = t1
        fu._feature_cache.put(det, seen["kp"], kp, des); t0 = tick("feature cache put (D2D)", t0)
        cur = (kp, des)
        c0 = fu._feature_cache.get(prev[0], prev[1], mat.device); c1 = fu._feature_cache.get(cur[0], cur[1], mat.device); t0 = tick("feature cache lookup x2", t0)
        assert c0 is not None and c1 is not None
        raw = mat.match_mixed(c0[0], c0[1], c1[0], c1[1]); t0 = tick("match_mixed (enqueue + sync + D2H)", t0)
        keep = raw["scores"] > np.float32(0.7); ms = fu._convert_lg_matches_to_opencv(raw["matches"][keep]); t0 = tick("DMatch list", t0)
        prev = cur
for k, v in T.items():
    print(f"{k:42s} {v / 10 * 1e3:7.3f} ms/pair")
print(f"{'total':42s} {sum(T.values()) / 10 * 1e3:7.3f} ms/pair   matches {len(ms)}")
# the same through the two public calls
torch.cuda.synchronize(); t0 = time.perf_counter()
for t in range(1, 11):
    cur = fu.feature_extractor(ns, frames[t], det); ms = fu.feature_matcher(ns, prev[0], cur[0], prev[1], cur[1], mat); prev = cur
print(f"feature_extractor + feature_matcher: {(time.perf_counter() - t0) * 100:.3f} ms/pair")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0 = fu._feature_cache.get(prev[0], prev[1], mat.device)
e0.record()
for _ in range(10):
    mat.match_device(c0[0], c0[1], c0[0], c0[1], full=False)
e1.record(); torch.cuda.synchronize()
print(f"match_device GPU time: {e0.elapsed_time(e1) / 10:.3f} ms/pair")
