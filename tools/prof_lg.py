"""LightGlue-only loop for ncu launch lists (run under gpurun): python tools/prof_lg.py [precision] [n_iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from b200slam import weights, frontend
from helpers import noisy_copy_pair

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
sl, _ = weights.load_lightglue_state(allow_synthetic=True)
mat = frontend.LightGlue(weights=sl, device=dev, precision=prec, max_kp=2048)
k0, d0, k1, d1, _ = noisy_copy_pair(2048, 2048, seed=1)
k0, d0, k1, d1 = k0.to(dev), d0.to(dev), k1.to(dev), d1.to(dev)
for _ in range(n):
    r = mat.match_device(k0, d0, k1, d1, full=False)
torch.cuda.synchronize()
print("matches", int(r["n"]), "stop", int(r["stop"]))
