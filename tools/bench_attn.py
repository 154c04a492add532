"""Device-only timing of one attention launch of a LightGlue self block at 2048 x 2048 (2 problems x 4 heads), fp32-faithful
(bf16x3) and bf16 kernels: python tools/bench_attn.py"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200slam._lib import lib, check
for name in ("b2s_bench_attn_tc3", "b2s_bench_attn_tc"):
    ms = C.c_float(0)
    check(getattr(lib, name)(2048, 2048, 200, C.addressof(ms)), name)
    fl = 2 * 4 * 4 * 2048 * 2048 * 64        # executed: 2 problems x 4 heads x (QK^T + PV)
    print(f"{name}: {ms.value * 1e3:.2f} us/launch  executed {fl / ms.value / 1e9:.1f} TFLOP/s  (B2S_ATTN_PV_ISSUERS={os.environ.get('B2S_ATTN_PV_ISSUERS', '2')})")
