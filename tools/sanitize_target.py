"""Small workloads for compute-sanitizer (run under gpurun): exercises every hand-written mbarrier / TMEM / TMA
pipeline at sizes the sanitizer finishes in minutes.
  python tools/sanitize_target.py [gemm|attn|conv|match|batch|geom|all]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from b200slam import _lib, frontend, synth, weights  # noqa: E402
from helpers import noisy_copy_pair  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
lib, check = _lib.lib, _lib.check
g = torch.Generator().manual_seed(0)


def gemm():
    for fn in ("b2s_test_gemm_tc", "b2s_test_gemm_tc3", "b2s_test_gemm_h2"):
        for (M, N, K) in [(200, 256, 256), (333, 512, 512), (130, 768, 256)]:
            A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
            out = np.empty((M, N), np.float32)
            check(getattr(lib, fn)(A.numpy().ctypes.data, W.numpy().ctypes.data, b.numpy().ctypes.data, M, N, K, out.ctypes.data), fn)
            ref = (A.double() @ W.double().T + b.double()).numpy()
            print(fn, M, N, K, "rel err", float(np.abs(out - ref).max() / np.abs(ref).max()))


def attn():
    for fn in ("b2s_test_attn_tc", "b2s_test_attn_tc3", "b2s_test_attn_h2"):
        for (nq, nk) in [(130, 200), (300, 390)]:
            q = torch.randn(nq, 256, generator=g); k = torch.randn(nk, 256, generator=g); v = torch.randn(nk, 256, generator=g)
            out = np.empty((nq, 256), np.float32)
            check(getattr(lib, fn)(q.numpy().ctypes.data, k.numpy().ctypes.data, v.numpy().ctypes.data, nq, nk, out.ctypes.data), fn)
            print(fn, nq, nk, "finite", bool(np.isfinite(out).all()))


def conv_and_match(batch=False):
    Hh, Ww, nkp = 120, 160, 256
    det = frontend.ALIKED(max_num_keypoints=nkp, weights=weights.synthetic_aliked_state(), device="cuda:0")
    f0, f1 = det.extract_bgr(synth.frame(0, Hh, Ww)), det.extract_bgr(synth.frame(1, Hh, Ww))
    print("extract", f0["keypoints"].shape[1], f1["keypoints"].shape[1])
    for prec in ("fp32", "bf16"):
        mat = frontend.LightGlue(weights=weights.synthetic_lightglue_state(token_bias=1.4, token_gain=6.0, match_bias=-5.0, match_gain=4.0),
                                 device="cuda:0", precision=prec, max_kp=512, filter_threshold=1e-6)
        r = mat({"image0": f0, "image1": f1})
        k0, d0, k1, d1, _ = noisy_copy_pair(300, 280, seed=1)
        r2 = mat.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy(), full=True)
        print(prec, "matches", len(r["matches"][0]), len(r2["matches"]), "stop", r2["stop"])
        if batch:
            feats = [noisy_copy_pair(200 + 40 * i, 200 + 40 * i, seed=20 + i)[:2] for i in range(3)]
            out = mat.match_batch_device([f[0].cuda() for f in feats], [f[1].cuda() for f in feats], [(0, 1), (0, 2), (1, 2), (2, 0)])
            torch.cuda.synchronize()
            print(prec, "batch", [int(v) for v in out["n"].cpu()], "stop", [int(v) for v in out["stop"].cpu()])
    if batch:
        imgs = [torch.from_numpy(synth.frame(t, Hh, Ww)).cuda() for t in range(3)]
        kp, de, sc, n = det.extract_batch_device(imgs, _lib.IMG_BGR_U8_HWC, Hh, Ww, 3 * Ww, lanes=2)
        torch.cuda.synchronize()
        print("extract_batch", n.cpu().tolist())


def geom():
    from b200slam.geometry import FundamentalRansac
    rng = np.random.default_rng(0)
    r = FundamentalRansac(max_points=1024, n_hyp=1024)
    p1 = rng.uniform(0, 600, (300, 2)).astype(np.float32)
    p2 = (p1 + rng.normal(0, 1.5, p1.shape) + np.array([12.0, 1.0])).astype(np.float32)
    F, mask = r.run_cv(p1, p2, 1.0, 0.99)
    F2, mask2 = r.run_host(p1, p2, 1.0)
    print("fm cv inliers", r.last_count, "info", r.last_info, "parallel", int(mask2.sum()))


if what in ("geom", "all"):
    geom()
if what in ("gemm", "all"):
    gemm()
if what in ("attn", "all"):
    attn()
if what in ("match", "conv", "all"):
    conv_and_match(batch=False)
if what in ("batch",):
    conv_and_match(batch=True)
torch.cuda.synchronize()
print("sanitize target done:", what)
