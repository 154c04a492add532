"""Batched LightGlue loop for ncu launch lists / in-library class timing (run under gpurun):
  python tools/prof_batch.py [precision=fp32] [B=8] [iters=2] [classes=0|1]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from b200slam import weights, frontend
from helpers import noisy_copy_pair

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
classes = len(sys.argv) > 4 and sys.argv[4] == "1"
dev = torch.device("cuda", 0)
sl, _ = weights.load_lightglue_state(allow_synthetic=True)
mat = frontend.LightGlue(weights=sl, device=dev, precision=prec, max_kp=2048)
feats = [noisy_copy_pair(2048, 2048, seed=1 + i)[:2] for i in range(3)]
kp = torch.cat([f[0] for f in feats]).to(dev); de = torch.cat([f[1] for f in feats]).to(dev)
cu = np.arange(4, dtype=np.int32) * 2048
pi = (np.arange(B, dtype=np.int32) % 2); pj = pi + 1
mat.reserve(2048, B)
out = mat.match_batch_packed(kp, de, cu, pi, pj, stride=2048)
torch.cuda.synchronize()
if classes:
    mat.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    mat.match_batch_packed(kp, de, cu, pi, pj, stride=2048, out=out)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / iters
msg = f"{prec} B={B}: {tot / B:.3f} ms/pair"
if classes:
    a, na = mat.profile_read(0); g, ng = mat.profile_read(1)
    msg += f" | attention {a / iters / B:.3f} ms/pair ({na // iters} launches/batch, {1e3 * a / na:.1f} us each) | gemm {g / iters / B:.3f} ms/pair ({ng // iters} launches/batch) | other {(tot - (a + g) / iters) / B:.3f}"
print(msg, "matches", out["n"][:B].cpu().tolist()[:3])
