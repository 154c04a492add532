"""ALIKED-only loop for ncu launch lists (run under gpurun): python tools/prof_aliked.py [n_frames]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib, weights, frontend, synth

H, W, NKP = 376, 1241, 2048
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
sa, _ = weights.load_aliked_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
frames = [torch.from_numpy(synth.frame(t, H, W)).to(dev) for t in range(n)]
for f in frames:
    det.extract_device(f, _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
torch.cuda.synchronize()
