"""Decision margins of the discrete outputs, run under gpurun: a flip of one of these decisions between two
correct fp32 implementations is only possible where the margin is at fp32-noise level (< 1e-6).
  detector : min |score - 0.2| over the NMS maxima, gap between the k-th and (k+1)-th candidate score (top-k cut)
  matcher  : per layer |exit ratio - 0.95|, min |match score - filter_threshold|, min |score - min_conf|
  python tools/report_margins.py [n_frames=4] > profiles/r2_decision_margins.json"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from b200slam import weights, frontend, synth
from helpers import aliked_selection

H, W, NKP = 376, 1241, 2048
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sa, _ = weights.load_aliked_state(allow_synthetic=True); sl, _ = weights.load_lightglue_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device="cuda:0")
mat = frontend.LightGlue(weights=sl, device="cuda:0", max_kp=NKP); mat.set_debug(True)
out = {"detector": [], "matcher": []}
feats = []
for t in range(n_frames):
    kp, de, sc = det.extract_host(synth.frame(t, H, W), desc_renorm_eps=1e-8)
    Hr, Wr = [int(v) for v in det.debug("geometry")][:2]
    S = det.debug("score_map").reshape(Hr, Wr)
    sel, nms, kth = aliked_selection(S, NKP)
    cand = np.sort(S.ravel()[np.flatnonzero(nms.ravel() > 0.2)])[::-1]
    out["detector"].append({"frame": t, "candidates_above_thr": int(len(cand)), "min_abs_score_minus_thr": float(np.abs(nms[nms > 0] - 0.2).min()),
                            "kth_minus_next": float(cand[NKP - 1] - cand[NKP]) if len(cand) > NKP else None,
                            "ties_at_cut": int((cand == cand[NKP - 1]).sum()) if len(cand) > NKP else 0})
    feats.append((kp, de))
for t in range(1, n_frames):
    r = mat.match_host(feats[t - 1][0], feats[t - 1][1], feats[t][0], feats[t][1], full=True)
    c = mat.debug("ctrl")
    npts = len(feats[t - 1][0]) + len(feats[t][0])
    ratios = [1.0 - float(c[8 + i]) / npts for i in range(r["stop"] - 1 if r["stop"] < mat.n_layers else mat.n_layers - 1)]
    ms0 = r["matching_scores0"]
    pos = ms0[ms0 > 0]
    out["matcher"].append({"pair": [t - 1, t], "stop": r["stop"], "exit_ratio_per_layer": ratios,
                           "min_abs_exit_ratio_minus_0.95": float(min(abs(x - 0.95) for x in ratios)) if ratios else None,
                           "min_abs_score_minus_filter_0.1": float(np.abs(pos - 0.1).min()) if len(pos) else None,
                           "min_abs_score_minus_min_conf_0.7": float(np.abs(pos - 0.7).min()) if len(pos) else None,
                           "matches": int(len(r["matches"]))})
print(json.dumps(out, indent=1))
