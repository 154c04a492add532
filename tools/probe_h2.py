"""fp16x2 vs bf16x3 operand schemes (run under gpurun): GEMM error vs float64 and per-launch time of the layer GEMM
shapes, attention kernel time.  python tools/probe_h2.py"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib
lib = _lib.lib
for (M, N, K) in [(4096, 768, 256), (4096, 512, 512), (4096, 256, 512)]:
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    ref = (A.double() @ W.double().T + b.double()).numpy()
    for name in ("b2s_test_gemm_tc3", "b2s_test_gemm_h2"):
        out = np.empty((M, N), np.float32)
        _lib.check(getattr(lib, name)(A.numpy().ctypes.data, W.numpy().ctypes.data, b.numpy().ctypes.data, M, N, K, out.ctypes.data), name)
        e = np.abs(out - ref)
        print(f"{name} {M}x{N}x{K}: max err {e.max() / np.abs(ref).max():.3e} rms {np.sqrt((e ** 2).mean()) / np.abs(ref).std():.3e} "
              f"(torch fp32 max {np.abs((A @ W.T + b).numpy() - ref).max() / np.abs(ref).max():.3e})", flush=True)
for (M, N, K) in [(32768, 768, 256), (32768, 512, 512), (32768, 256, 512), (32768, 512, 256), (4096, 768, 256), (4096, 512, 512)]:
    for name in ("b2s_bench_gemm_tc3", "b2s_bench_gemm_h2"):
        ms = C.c_float(0)
        _lib.check(getattr(lib, name)(M, N, K, 0, 20, C.addressof(ms), None, None), name)
        print(f"{name} persistent {M}x{N}x{K}: {ms.value * 1e3:.1f} us  {2.0 * M * N * K / ms.value / 1e9:.1f} TFLOP/s (algorithmic)", flush=True)
for (nq, nk) in [(2048, 2048), (1024, 1024)]:
    for name in ("b2s_bench_attn_tc3", "b2s_bench_attn_h2", "b2s_bench_attn_tc"):
        ms = C.c_float(0)
        _lib.check(getattr(lib, name)(nq, nk, 20, C.addressof(ms)), name)
        print(f"{name} {nq}x{nk}: {ms.value * 1e3:.1f} us  {2 * 4 * nq * nk * 64 * 2 * 2 / ms.value / 1e9:.1f} TFLOP/s (algorithmic)", flush=True)
