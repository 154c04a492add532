"""Extraction-only throughput of b2s_aliked_extract_batch (run under gpurun): ms per frame as a function of the number of
concurrent extractor lanes, GPU otherwise idle.  python tools/time_extract_batch.py [B=8]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib, weights, frontend, synth

H, W, NKP = 376, 1241, 2048
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
sa, _ = weights.load_aliked_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
frames = [torch.from_numpy(synth.frame(t, H, W)).to(dev) for t in range(B)]
out = None
for lanes in (1, 2, 3, 4, 6, 8):
    for _ in range(3):
        out = det.extract_batch_device(frames, _lib.IMG_BGR_U8_HWC, H, W, 3 * W, lanes=lanes, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10):
        out = det.extract_batch_device(frames, _lib.IMG_BGR_U8_HWC, H, W, 3 * W, lanes=lanes, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} lanes={lanes}: {e0.elapsed_time(e1) / 10 / B:.3f} ms/frame (wall {(time.perf_counter() - t0) * 100 / B:.3f}) n={out[3].cpu().tolist()[:3]}", flush=True)
