"""Device-only kernel timings (run under gpurun):  python tools/bench_kernels.py attn"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import _lib

def attn():
    for nq, nk in [(2048, 2048), (1536, 1536), (1024, 1024), (512, 512)]:
        ms = C.c_float(0)
        _lib.check(_lib.lib.b2s_bench_attn_tc(nq, nk, 50, C.byref(ms)), "bench_attn")
        fl = 2 * 4 * 4.0 * nq * nk * 64          # executed: QK^T + PV, 4 heads, 2 problems
        print(f"attn_tc {nq}x{nk}: {ms.value*1e3:.2f} us/launch  {fl/ms.value/1e9:.1f} TFLOP/s executed", flush=True)
        _lib.check(_lib.lib.b2s_bench_attn_tc3(nq, nk, 50, C.byref(ms)), "bench_attn3")
        print(f"attn_tc3 {nq}x{nk}: {ms.value*1e3:.2f} us/launch  {fl/ms.value/1e9:.1f} TFLOP/s fp32-equivalent ({6*fl/ms.value/1e9:.1f} bf16 issued)", flush=True)

if __name__ == "__main__":
    {"attn": attn}[sys.argv[1]]()
