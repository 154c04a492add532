"""Single-stream device timings (run under gpurun): ALIKED extract, LightGlue match as a function of the pairs per launch
sequence (b2s_lightglue_match_batch_ex), real extracted features of consecutive synthetic KITTI-shaped frames.
  python tools/time_batch.py [precisions=fp32,bf16] [batches=1,2,4,8,16] [nkp=2048]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib, weights, frontend, synth

H, W = 376, 1241
precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "bf16"]
batches = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 16]
NKP = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
dev = torch.device("cuda", 0)
sa, _ = weights.load_aliked_state(allow_synthetic=True); sl, _ = weights.load_lightglue_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
NF = 17
frames = [torch.from_numpy(synth.frame(t, H, W)).to(dev) for t in range(NF)]
kps, des = [], []
for f in frames:
    kp, de, sc, n = det.extract_device(f, _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
    torch.cuda.synchronize()
    k = int(n.item()); kps.append(kp[:k].clone()); des.append(de[:k].clone())
cu = np.cumsum([0] + [len(k) for k in kps]).astype(np.int32)
KP, DE = torch.cat(kps), torch.cat(des)


def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.perf_counter() - t0) * 1e3 / iters


ms, wall = timeit(lambda: det.extract_device(frames[1], _lib.IMG_BGR_U8_HWC, H, W, 3 * W), 20)
print(f"ALIKED extract: {ms:.3f} ms/frame (wall {wall:.3f})  kp={len(kps[1])}", flush=True)
for prec in precs:
    for dc, wc, tag in ((0.95, 0.99, "adaptive"), (-1, -1, "full-depth")):
        mat = frontend.LightGlue(weights=sl, device=dev, precision=prec, max_kp=NKP, depth_confidence=dc, width_confidence=wc)
        for B in batches:
            pi = np.arange(0, B, dtype=np.int32) % (NF - 1); pj = pi + 1
            mat.reserve(NKP, B)
            out = mat.match_batch_packed(KP, DE, cu, pi, pj, stride=NKP)
            l0 = mat.launches
            ms, wall = timeit(lambda: mat.match_batch_packed(KP, DE, cu, pi, pj, stride=NKP, out=out), 10)
            nl = (mat.launches - l0) / 13
            print(f"LightGlue {prec} {tag} B={B:2d}: {ms / B:.3f} ms/pair ({ms:.3f} ms/batch, wall {wall:.3f}) -> {1e3 * B / ms:7.1f} pairs/s | "
                  f"launches/pair {nl / B:.1f} | matches {out['n'][:B].cpu().tolist()[:4]} stop {out['stop'][:B].cpu().tolist()[:4]}", flush=True)
        del mat
