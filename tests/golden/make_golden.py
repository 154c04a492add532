"""Generates tests/golden/aliked_lg_small.npz + meta.json from the CPU oracle on seeded
synthetic inputs (no reference import is possible: `lightglue`/`kornia` are not installed and
the reference holds no fixtures for this path - SURVEY.md 8c).  Run: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from b200slam import weights, synth  # noqa: E402

META = {"model": "aliked-n16", "seed": 0, "max_kp": 512, "H": 240, "W": 320, "min_matches": 100}


@torch.no_grad()
def main():
    det = oracle.ALIKED(model_name=META["model"], max_num_keypoints=META["max_kp"]).eval()
    det.load_state_dict(weights.synthetic_aliked_state(META["model"], META["seed"]), strict=True)
    mat = oracle.LightGlue().eval()
    mat.load_state_dict(weights.synthetic_lightglue_state(META["seed"]), strict=False)
    H, W = META["H"], META["W"]
    f0 = det.extract(oracle.bgr_to_tensor(synth.frame(0, H, W)))
    f1 = det.extract(oracle.bgr_to_tensor(synth.frame(1, H, W)))
    r = mat({"image0": f0, "image1": f1})
    out = dict(kp0=f0["keypoints"][0].numpy(), desc0=f0["descriptors"][0].numpy().astype(np.float32),
               kp1=f1["keypoints"][0].numpy(), desc1=f1["descriptors"][0].numpy().astype(np.float32),
               matches=r["matches"][0].numpy().astype(np.int32), scores=r["scores"][0].numpy(),
               stop=np.int32(r["stop"]))
    np.savez_compressed(os.path.join(HERE, "aliked_lg_small.npz"), **out)
    json.dump(META, open(os.path.join(HERE, "meta.json"), "w"), indent=1)
    print("matches", len(out["matches"]), "kp", len(out["kp0"]), len(out["kp1"]), "stop", r["stop"])


if __name__ == "__main__":
    main()
