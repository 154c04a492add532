"""Host-side breakdown of the drop-in API (run under gpurun)."""
import os, sys, time
from types import SimpleNamespace
import numpy as np, torch, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import features_utils as fu, synth, frontend, weights

H, W, NKP = 376, 1241, 2048
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ns = SimpleNamespace(use_lightglue=True, max_features=NKP, min_conf=0.7, lg_precision=prec)
sa, _ = weights.load_aliked_state(allow_synthetic=True); sl, _ = weights.load_lightglue_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device="cuda:0")
mat = frontend.LightGlue(weights=sl, device="cuda:0", precision=prec, max_kp=NKP)
frames = [synth.frame(t, H, W) for t in range(12)]
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
prev = fu.feature_extractor(ns, frames[0], det)
for rep in range(2):
    T.clear()
    for t in range(1, 11):
        t0 = time.perf_counter(); kps, des, _ = det.extract_host(frames[t]); tick("extract_host (C call)", t0)
        t0 = time.perf_counter(); kp = fu._convert_lg_kps_to_opencv(kps); tick("KeyPoint list", t0)
        t0 = time.perf_counter(); des = des.astype(np.float32, copy=True); des /= (np.linalg.norm(des, axis=1, keepdims=True) + 1e-8).astype(np.float32); tick("desc renorm", t0)
        cur = (kp, des)
        t0 = time.perf_counter(); a0 = fu._kps_to_array(prev[0]); a1 = fu._kps_to_array(cur[0]); tick("KeyPoint->array x2", t0)
        t0 = time.perf_counter(); raw = mat.match_host(a0, prev[1], a1, cur[1]); tick("match_host (C call)", t0)
        t0 = time.perf_counter(); keep = raw["scores"] > np.float32(0.7); ms = fu._convert_lg_matches_to_opencv(raw["matches"][keep]); tick("DMatch list", t0)
        prev = cur
for k, v in T.items():
    print(f"{k:28s} {v / 10 * 1e3:7.3f} ms/pair")
print(f"{'total':28s} {sum(T.values()) / 10 * 1e3:7.3f} ms/pair   matches {len(ms)}")
