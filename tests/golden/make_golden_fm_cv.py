"""Generates tests/golden/fm_cv.npz by running the reference's actual implementation of `filter_matches_ransac`'s core -
`cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, 0.99)` (/root/reference/slam/core/features_utils.py:195-196;
OpenCV is the reference's un-vendored dependency, opencv-python-headless 4.13.0.92 in this image) - on the seeded
two-view scenes of oracle/geometry.py::two_view_scene.     python tests/golden/make_golden_fm_cv.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cv_ransac as R               # noqa: E402
from oracle import geometry as G                # noqa: E402

#        n, outlier fraction, noise px, seed, thresh, integer grid
CASES = [(700, 0.25, 0.4, 11, 1.0, False), (1300, 0.40, 0.5, 12, 1.0, False), (2048, 0.55, 0.5, 13, 1.0, False),
         (300, 0.75, 0.8, 14, 2.5, False), (15, 0.2, 0.3, 15, 1.0, False), (600, 0.3, 0.5, 16, 2.0, True),
         (4000, 0.35, 0.6, 17, 3.0, False), (64, 0.5, 0.3, 18, 0.5, False)]


def scene(n, of, noise, seed, grid):
    p1, p2, _ = G.two_view_scene(n, of, noise, seed=seed)
    if grid:        # ORB-like integer coordinates with duplicated points: exercises getSubset's redraws / checkSubset
        p1, p2 = np.rint(p1).astype(np.float32), np.rint(p2).astype(np.float32)
        p1[::7] = p1[0]
    return p1, p2


if __name__ == "__main__":
    out = {"cv2_version": np.array(cv2.__version__)}
    for c, (n, of, noise, seed, thresh, grid) in enumerate(CASES):
        p1, p2 = scene(n, of, noise, seed, grid)
        F, mask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, thresh, 0.99)
        Fo, mo = R.find_fundamental_mat(p1, p2, thresh, 0.99)
        same = np.array_equal(mask, mo) and np.abs(F - Fo).max() <= 1e-9 * np.abs(F).max()
        print(c, (n, of, noise, seed, thresh, grid), "cv2 inliers", int(mask.sum()), "restatement identical:", same)
        out[f"c{c}_cfg"] = np.array([n, of, noise, seed, thresh, float(grid)])
        out[f"c{c}_pts1"], out[f"c{c}_pts2"] = p1, p2
        out[f"c{c}_F"] = F
        out[f"c{c}_mask"] = np.packbits(mask.ravel())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fm_cv.npz"), **out)
