"""Per-kernel roofline table from the ncu raw pages committed under profiles/ (ncu --set full --clock-control none,
exported with `ncu -i X.ncu-rep --page raw --csv`).  For every distinct kernel (name + grid) of the captures: ncu launch
duration (cold caches, serialised), DRAM bytes, achieved DRAM GB/s and its fraction of the measured HBM copy bandwidth
(MEASURED_PEAKS.json), tensor-pipe active %, registers.  Usage: python tools/roofline_table.py profiles/r1_ncu_*.raw.csv"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6553.3))


def col(hdr, name):
    return hdr.index(name) if name in hdr else None


agg = OrderedDict()
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    c = {k: col(hdr, k) for k in ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                   "launch__registers_per_thread", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                                   "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                                   "sm__warps_active.avg.pct_of_peak_sustained_active")}
    units = rows[1]

    def f(r, k, unit_scale=None):
        i = c[k]
        if i is None or r[i] in ("", "no data", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        if unit_scale:
            u = units[i]
            v *= unit_scale.get(u, 1.0)
        return v

    for r in rows[2:]:
        if len(r) <= c["Kernel Name"]:
            continue
        name = r[c["Kernel Name"]].split("(")[0].replace("void ", "").replace("b2s::", "")
        key = (name, r[c["Grid Size"]], r[c["Block Size"]])
        byte = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        t = f(r, "gpu__time_duration.sum", {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3})
        rd, wr = f(r, "dram__bytes_read.sum", byte) or 0.0, f(r, "dram__bytes_write.sum", byte) or 0.0
        e = agg.setdefault(key, {"n": 0, "us": 0.0, "bytes": 0.0, "tensor": 0.0, "sm": 0.0, "lts": 0.0, "warps": 0.0, "regs": f(r, "launch__registers_per_thread")})
        e["n"] += 1; e["us"] += t or 0.0; e["bytes"] += rd + wr
        e["tensor"] += f(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active") or 0.0
        e["sm"] += f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0.0
        e["lts"] += f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed") or 0.0
        e["warps"] += f(r, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0.0

print(f"HBM denominator: {HBM:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs); durations are ncu launch times (cold caches, serialised)\n")
print("| kernel | grid | block | regs | launches | avg us | DRAM MB / launch | DRAM GB/s | % HBM peak | tensor pipe active % | SM throughput % | L2 throughput % | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for (name, grid, block), e in agg.items():
    n = e["n"]; us = e["us"] / n; mb = e["bytes"] / n / 1e6
    gbs = e["bytes"] / n / (us * 1e-6) / 1e9 if us > 0 else 0.0
    print(f"| `{name}` | {grid} | {block} | {int(e['regs'] or 0)} | {n} | {us:.1f} | {mb:.2f} | {gbs:.0f} | {100 * gbs / HBM:.1f} | {e['tensor'] / n:.1f} | {e['sm'] / n:.1f} | {e['lts'] / n:.1f} | {e['warps'] / n:.1f} |")
