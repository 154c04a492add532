"""Per-phase timing of the bf16x3 layer GEMM (b2s_bench_gemm_tc3): mean launch time back to back, and where a CTA's
time goes (globaltimer stamps: start -> deps -> first stage -> accumulators -> epilogue -> end)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from b200slam._lib import lib, check   # noqa: E402

for (M, N, K, cl) in [(4096, 512, 256, 1), (4096, 512, 256, 4), (4096, 512, 512, 1), (4096, 512, 512, 4), (4096, 768, 256, 1),
                      (4096, 768, 256, 2), (4096, 256, 512, 1), (4096, 256, 512, 4)]:
    ms = C.c_float(0); n = C.c_int(0)
    ts = np.zeros((4096 * 6,), np.int64)
    check(lib.b2s_bench_gemm_tc3(M, N, K, cl, 200, C.addressof(ms), ts.ctypes.data, C.addressof(n)), "bench_gemm")
    t = ts[:n.value * 6].reshape(-1, 6).astype(np.float64) / 1e3
    d = np.diff(t, axis=1)
    print(f"M={M} N={N} K={K} cl={cl}: {ms.value * 1e3:7.2f} us/launch | CTAs {n.value} | start spread {t[:, 0].max():5.2f} us | "
          f"deps {d[:, 0].mean():5.2f} first-stage {d[:, 1].mean():5.2f} mainloop {d[:, 2].mean():5.2f} epilogue {d[:, 3].mean():5.2f} "
          f"teardown {d[:, 4].mean():5.2f} | last CTA end {t[:, 5].max():5.2f} us")
