"""Probe the tensor-core accumulator rounding through the bf16x3 GEMM (run under gpurun)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib
for K in (64, 256, 1024, 2048, 4096):
    for kind in ("positive", "signed"):
        g = torch.Generator().manual_seed(K)
        M, N = 128, 64
        A = torch.rand(M, K, generator=g) if kind == "positive" else torch.randn(M, K, generator=g)
        W = torch.rand(N, K, generator=g) if kind == "positive" else torch.randn(N, K, generator=g)
        out = np.empty((M, N), np.float32)
        _lib.check(_lib.lib.b2s_test_gemm_tc3(A.numpy().ctypes.data, W.numpy().ctypes.data, None, M, N, K, out.ctypes.data), "gemm")
        ref = (A.double() @ W.double().T).numpy()
        t32 = (A @ W.T).numpy()
        scale = np.abs(ref).max()
        e = (out - ref) / scale; e32 = (t32 - ref) / scale
        print(f"K={K:5d} {kind:8s} tc3: mean {e.mean():+.2e} max {np.abs(e).max():.2e} | torch fp32: mean {e32.mean():+.2e} max {np.abs(e32).max():.2e}")
