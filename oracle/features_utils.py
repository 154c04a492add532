"""Oracle: the reference adapter `slam/core/features_utils.py` restated over the CPU oracle
models (TEST INFRASTRUCTURE - see oracle/__init__.py).  Same function names, argument
meaning and return types as the reference (LightGlue branch only):

  init_feature_pipeline   features_utils.py:18-30
  feature_extractor       features_utils.py:85-101
  feature_matcher         features_utils.py:109-171
  _lightglue_detect_and_match   features_utils.py:233-247

This is what bench.py times as the CPU baseline (through list[cv2.KeyPoint]/list[cv2.DMatch]).
"""
import cv2
import numpy as np
import torch

from .aliked import ALIKED
from .lightglue import LightGlue
from .preprocess import bgr_to_tensor
from .utils import rbd


def init_feature_pipeline(args, aliked_state=None, lightglue_state=None):
    detector = ALIKED(max_num_keypoints=int(getattr(args, "max_features", 4000))).eval()
    matcher = LightGlue(features="aliked").eval()
    if aliked_state is not None:
        detector.load_state_dict(aliked_state, strict=True)
    if lightglue_state is not None:
        matcher.load_state_dict(lightglue_state, strict=False)
    return detector, matcher


def feature_extractor(args, img, detector):
    t0 = bgr_to_tensor(img)
    feats = rbd(detector.extract(t0))
    kp0 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in feats["keypoints"]]
    des0 = feats["descriptors"].detach().to("cpu").float().numpy()
    des0 /= (np.linalg.norm(des0, axis=1, keepdims=True) + 1e-8).astype(np.float32)
    return kp0, des0


def feature_matcher(args, kp0, kp1, des0, des1, matcher):
    if (des0 is None or des1 is None or kp0 is None or kp1 is None
            or len(kp0) == 0 or len(kp1) == 0 or len(des0) == 0 or len(des1) == 0):
        return []
    k0 = torch.tensor([(k.pt[0], k.pt[1]) for k in kp0], dtype=torch.float32)[None]
    k1 = torch.tensor([(k.pt[0], k.pt[1]) for k in kp1], dtype=torch.float32)[None]
    d0 = torch.as_tensor(des0, dtype=torch.float32)[None]
    d1 = torch.as_tensor(des1, dtype=torch.float32)[None]
    with torch.inference_mode():
        raw = matcher({"image0": {"keypoints": k0, "descriptors": d0},
                       "image1": {"keypoints": k1, "descriptors": d1}})
    raw = rbd(raw)
    matches_raw = raw["matches"]
    thr = float(getattr(args, "min_conf", 0.7))
    matches_raw = matches_raw[raw["scores"] > thr]
    return [cv2.DMatch(int(i), int(j), 0, 0.0) for i, j in matches_raw.cpu().numpy().tolist()]


def _lightglue_detect_and_match(img1, img2, extractor, matcher):
    f0, f1 = extractor.extract(bgr_to_tensor(img1)), extractor.extract(bgr_to_tensor(img2))
    matches = rbd(matcher({"image0": f0, "image1": f1}))
    f0, f1 = rbd(f0), rbd(f1)
    cv_kp0 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in f0["keypoints"]]
    cv_kp1 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in f1["keypoints"]]
    cv_m = [cv2.DMatch(int(i), int(j), 0, 0.0) for i, j in matches["matches"]]
    return cv_kp0, cv_kp1, f0["descriptors"], f1["descriptors"], cv_m
