"""Timeline of one CTA of the fp32 attention kernel from SM-clock stamps: python tools/trace_attn.py [h2|x3] (fp16x2 planes, default, or bf16x3)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from b200slam._lib import lib, check   # noqa: E402

trace_fn = lib.b2s_trace_attn_tc3 if (len(sys.argv) > 1 and sys.argv[1] == "x3") else lib.b2s_trace_attn_h2
ms = C.c_float(0)
for cta in ((0, 0, 0), (15, 3, 1), (7, 2, 0)):       # whole-CTA phases of a few CTAs (tr[0,0,0] carries the CTA to trace)
    tr = np.zeros((3, 64, 8), np.int64)
    tr[0, 0, 0] = cta[0] | (cta[1] << 8) | (cta[2] << 16)
    check(trace_fn(2048, 2048, 20, C.addressof(ms), tr.ctypes.data), "trace_attn")
    e = tr[0, 63]
    print(f"CTA {cta}: launch {ms.value * 1e3:.1f} us | entry -> deps {e[1] - e[0]} clk | deps -> key loop done {tr[1, 63, 0] - e[1]} | loop done -> output written {e[2] - tr[1, 63, 0]} | total {e[3] - e[0]} clk")
tr = np.zeros((3, 64, 8), np.int64)
check(trace_fn(2048, 2048, 50, C.addressof(ms), tr.ctypes.data), "trace_attn")
tr[0, 63] = 0; tr[1, 63] = 0; tr[2, 63] = 0
t0 = tr[tr > 0].min()
rel = np.where(tr > 0, tr - t0, -1)
print(f"launch {ms.value * 1e3:.1f} us; stamps in SM clocks relative to the first one")
print("pair |  MMA: K landed  S issued | PV0 start  issued | PV1 start  issued | s_free(next) ||  g0: S seen  exp done  PVprev done  P out ||  g1: S seen  exp done  PVprev done  P out")
for jp in range(16):
    m = rel[0, jp]; a = rel[1, jp]; b = rel[2, jp]
    print(f"{jp:4d} | {m[0]:8d} {m[1]:8d} | {m[2]:8d} {m[3]:8d} | {m[4]:8d} {m[5]:8d} | {m[6]:8d} || {a[0]:8d} {a[1]:8d} {a[2]:8d} {a[3]:8d} || {b[0]:8d} {b[1]:8d} {b[2]:8d} {b[3]:8d}")
d = np.diff(rel[1, 1:16, 3]); print("group 0 period (clk/tile):", d.mean().round(), " total clk:", rel.max())
a = rel[1, 2:15]
print("group 0 means: S seen -> exp done", (a[:, 1] - a[:, 0]).mean().round(), "| wait PVprev", (a[:, 2] - a[:, 1]).mean().round(), "| add_pv + store P", (a[:, 3] - a[:, 2]).mean().round(),
      "| P out -> next S seen", (rel[1, 3:16, 0] - rel[1, 2:15, 3]).mean().round())
