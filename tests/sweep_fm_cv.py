"""GPU sweep: b2s_fm_cv_ransac_host against cv2.findFundamentalMat(FM_RANSAC) over many seeded scenes; prints one JSON
line (scenes, identical masks, worst relative F difference, mean times).   python tests/sweep_fm_cv.py [scenes]"""
import json
import os
import sys
import time

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200slam.geometry import FundamentalRansac          # noqa: E402
from oracle.geometry import two_view_scene              # noqa: E402  (test infrastructure: scene generator)

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 300
r = FundamentalRansac(max_points=4096, n_hyp=2048)
sizes = [15, 20, 33, 64, 150, 300, 700, 1000, 1300, 2048, 3000, 4000]
same = 0
worst = 0.0
t_cv = t_gpu = 0.0
bad = []
per_size = {}
for s in range(n_scenes):
    n = sizes[s % len(sizes)]
    of, noise, thresh = [0.1, 0.3, 0.5, 0.7, 0.2][s % 5], [0.2, 0.5, 1.0][s % 3], [1.0, 1.0, 3.0, 0.5][s % 4]
    p1, p2, _ = two_view_scene(n, of, noise, seed=9000 + s)
    t0 = time.perf_counter()
    Fc, mc = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, thresh, 0.99)
    t1 = time.perf_counter()
    Fg, mg = r.run_cv(p1, p2, thresh, 0.99)
    t2 = time.perf_counter()
    t_cv += t1 - t0
    t_gpu += t2 - t1
    d = per_size.setdefault(n, [0.0, 0.0, 0])
    d[0] += t1 - t0; d[1] += t2 - t1; d[2] += 1
    ok = mg is not None and np.array_equal(mc, mg)
    same += ok
    if ok:
        worst = max(worst, float(np.abs(Fc - Fg).max() / np.abs(Fc).max()))
    else:
        bad.append([s, n, int(mc.sum()), None if mg is None else int(mg.sum())])
print(json.dumps({"scenes": n_scenes, "identical_masks": int(same), "worst_rel_F_diff": worst, "mismatches": bad,
                  "cv2_ms_mean": 1e3 * t_cv / n_scenes, "gpu_ms_mean": 1e3 * t_gpu / n_scenes,
                  "per_size_ms": {str(k): {"cv2": 1e3 * v[0] / v[2], "gpu": 1e3 * v[1] / v[2]} for k, v in sorted(per_size.items())},
                  "cv2_version": cv2.__version__}))
