"""Single-stream device timings of the two stages (run under gpurun): ALIKED extract and LightGlue match."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib, weights, frontend, synth

H, W, NKP = 376, 1241, 2048
dev = torch.device("cuda", 0)
sa, _ = weights.load_aliked_state(allow_synthetic=True); sl, _ = weights.load_lightglue_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
frames = [torch.from_numpy(synth.frame(t, H, W)).to(dev) for t in range(4)]
feats = []
for f in frames:
    kp, de, sc, n = det.extract_device(f, _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
    torch.cuda.synchronize()
    k = int(n.item()); feats.append((kp[:k].clone(), de[:k].clone()))
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.perf_counter() - t0) * 1e3 / iters
ms, wall = timeit(lambda: det.extract_device(frames[1], _lib.IMG_BGR_U8_HWC, H, W, 3 * W))
print(f"ALIKED extract: {ms:.3f} ms/frame (wall {wall:.3f})  kp={len(feats[1][0])}")
precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16", "fp32", "fp32x3"]
for prec in [p for p in precs if p != "none"]:
    for dc, wc, tag in ((0.95, 0.99, "adaptive"), (-1, -1, "full-depth")):
        mat = frontend.LightGlue(weights=sl, device=dev, precision=prec, max_kp=NKP, depth_confidence=dc, width_confidence=wc)
        r = mat.match_device(feats[0][0], feats[0][1], feats[1][0], feats[1][1], full=False)
        torch.cuda.synchronize()
        ms, wall = timeit(lambda: mat.match_device(feats[0][0], feats[0][1], feats[1][0], feats[1][1], full=False), 20)
        print(f"LightGlue {prec} {tag}: {ms:.3f} ms/pair (wall {wall:.3f}) matches={int(r['n'].item())} stop={int(r['stop'].item())}")
