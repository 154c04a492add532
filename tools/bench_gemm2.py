"""Layer-GEMM shapes (bf16x3 planes, fp32-faithful) back to back on one stream: the one-tile-per-CTA kernel (cl = 1) against
the persistent tile-scheduler kernel (cl = 0), for one pair (M = 4096 rows) and a batch of 8 pairs (M = 32768)."""
import ctypes as C
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from b200slam._lib import lib, check   # noqa: E402

for M in (4096, 32768):
    for (N, K) in [(768, 256), (512, 512), (256, 512), (512, 256)]:
        row = []
        for cl in (1, 0):
            ms = C.c_float(0); n = C.c_int(0)
            check(lib.b2s_bench_gemm_tc3(M, N, K, cl, 100 if M == 4096 else 30, C.addressof(ms), None, C.addressof(n)), "bench_gemm")
            row.append(ms.value * 1e3)
        fl = 2.0 * M * N * K
        print(f"M={M:6d} N={N} K={K}: one-tile-per-CTA {row[0]:7.2f} us | persistent {row[1]:7.2f} us ({row[0] / row[1]:.2f}x) | "
              f"persistent: {fl / row[1] / 1e6:6.1f} TFLOP/s algorithmic, {6 * fl / row[1] / 1e6:6.1f} issued", flush=True)
