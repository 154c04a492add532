"""SM-clock trace of CTA 0 of the persistent GEMM (run under gpurun): who waits for whom, per tile.
  python tools/trace_gemm.py [np=2] [M=32768]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib
np_ = int(sys.argv[1]) if len(sys.argv) > 1 else 2
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
fn = _lib.lib.b2s_bench_gemm_h2 if np_ == 2 else _lib.lib.b2s_bench_gemm_tc3
for (N, K) in [(768, 256), (512, 512), (256, 512), (512, 256)]:
    ms = C.c_float(0); ts = np.zeros(256, np.int64); nc = C.c_int(0)
    _lib.check(fn(M, N, K, 0, 10, C.addressof(ms), ts.ctypes.data, C.addressof(nc)), "bench_gemm")
    t = ts.reshape(4, 16, 4)
    t0 = t[1, 0, 0]
    print(f"--- np={np_} {M}x{N}x{K}: {ms.value * 1e3:.1f} us/launch; CTA 0 (cycles relative to its first tile)")
    print("tile | MMA: start  buf-free  first-stage  issued | EPI w2: wait  full  handback  stored | EPI w6: wait full handback stored | PROD: first-req last-slot-free")
    for i in range(1, 9):
        if t[1, i, 0] == 0: break
        r = lambda a: " ".join(f"{int(v - t0):7d}" for v in a)
        print(f"{i:4d} | {r(t[1, i])} | {r(t[2, i])} | {r(t[3, i])} | {r(t[0, i, :2])}")
    d = np.diff(t[1, 1:9, 3])
    print("MMA issue-complete to issue-complete per tile:", [int(v) for v in d if v > 0])
    print("epilogue w2 full->stored per tile:", [int(t[2, i, 3] - t[2, i, 1]) for i in range(1, 8) if t[2, i, 3] > 0])
    print("epilogue w6 full->stored per tile:", [int(t[3, i, 3] - t[3, i, 1]) for i in range(1, 8) if t[3, i, 3] > 0])
    print("MMA: wait for buffer:", [int(t[1, i, 1] - t[1, i, 0]) for i in range(1, 8)], "wait for first stage:", [int(t[1, i, 2] - t[1, i, 1]) for i in range(1, 8)],
          "main loop:", [int(t[1, i, 3] - t[1, i, 2]) for i in range(1, 8)])
