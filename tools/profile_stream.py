"""Where a chunk of FramePairStream goes (run under gpurun): host time of _enqueue / _collect per chunk next to the GPU time of
the chunk's kernels.  python tools/profile_stream.py [n_frames=129]"""
import os, sys, time
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import weights, frontend, synth, stream

H, W, NKP = 376, 1241, 2048
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 129
dev = torch.device("cuda", 0)
sa, _ = weights.load_aliked_state(allow_synthetic=True); sl, _ = weights.load_lightglue_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
mat = frontend.LightGlue(weights=sl, device=dev, max_kp=NKP)
frames = [synth.frame(t % 17, H, W) for t in range(n_frames)]
ns = SimpleNamespace(use_lightglue=True, min_conf=0.7)
fps = stream.FramePairStream(ns, det, mat)
for _ in fps.run(frames[:17]):
    pass
torch.cuda.synchronize()
t_enq, t_col = [], []
oe, oc = fps._enqueue, fps._collect
def enq(*a, **k):
    t0 = time.perf_counter(); r = oe(*a, **k); t_enq.append(time.perf_counter() - t0); return r
def col(*a, **k):
    t0 = time.perf_counter(); r = oc(*a, **k); t_col.append(time.perf_counter() - t0); return r
fps._enqueue, fps._collect = enq, col
t0 = time.perf_counter()
n = sum(1 for _ in fps.run(frames))
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"{n} frames in {tot * 1e3:.1f} ms -> {(n - 1) / tot:.1f} pairs/s | per chunk of 8: total {tot * 1e3 / (n / 8):.2f} ms, "
      f"_enqueue {np.mean(t_enq) * 1e3:.2f} ms (max {np.max(t_enq) * 1e3:.2f}), _collect {np.mean(t_col) * 1e3:.2f} ms (of which wait for D2H: see below)")
# split _collect: wait vs conversion
waits = []
oc2 = oc
def col2(c, nb, first, t0_, prev):
    k = c % 2
    a = time.perf_counter(); fps.ev_out[k].synchronize(); waits.append(time.perf_counter() - a)
    return oc2(c, nb, first, t0_, prev)
fps._collect = col2; fps._enqueue = oe
t0 = time.perf_counter(); n = sum(1 for _ in fps.run(frames)); torch.cuda.synchronize(); tot = time.perf_counter() - t0
print(f"second pass: {(n - 1) / tot:.1f} pairs/s | mean wait for the chunk's results inside _collect {np.mean(waits) * 1e3:.2f} ms per chunk "
      f"(0 = host-bound, > 0 = GPU-bound)")
