"""Generates tests/golden/pnp_reproj.npz by running the REFERENCE's own `reproject_and_match_2d3d`
(/root/reference/slam/core/pnp_utils.py:224-304, importable in the build container: numpy + cv2 + scipy only) on the
seeded scenes of oracle/pnp.py::tracking_scene.  Run here, never on the GPU box:  python tests/golden/make_golden_pnp.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from slam.core import pnp_utils as ref          # noqa: E402
from oracle import pnp as O                     # noqa: E402

CASES = [dict(n_points=1500, n_kps=2048, seed=0, radius=12.0, max_l2=0.8, cosine=False),
         dict(n_points=3000, n_kps=2048, seed=1, radius=20.0, max_l2=0.9, cosine=False),
         dict(n_points=400, n_kps=300, seed=2, radius=30.0, max_l2=0.7, cosine=False),
         dict(n_points=1500, n_kps=2048, seed=3, radius=12.0, max_l2=0.35, cosine=True)]

out = {}
for c, cfg in enumerate(CASES):
    wm, K, Tcw, kps, des = O.tracking_scene(cfg["n_points"], cfg["n_kps"], cfg["seed"])
    r = ref.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, radius_px=cfg["radius"], max_l2=cfg["max_l2"],
                                     use_cosine=cfg["cosine"])
    o = O.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, radius_px=cfg["radius"], max_l2=cfg["max_l2"],
                                   use_cosine=cfg["cosine"])
    print(c, cfg, "reference matches", len(r.mp_ids), "restatement identical:", r.mp_ids == o.mp_ids and r.kp_indices == o.kp_indices)
    out[f"c{c}_cfg"] = np.array([cfg["n_points"], cfg["n_kps"], cfg["seed"], cfg["radius"], cfg["max_l2"], float(cfg["cosine"])])
    out[f"c{c}_mp_ids"] = np.asarray(r.mp_ids, np.int64)
    out[f"c{c}_kp_indices"] = np.asarray(r.kp_indices, np.int64)
    out[f"c{c}_pts3d"] = r.pts3d
    out[f"c{c}_pts2d"] = r.pts2d
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pnp_reproj.npz"), **out)
