"""Per-kernel SASS opcode summary of libb200slam.so (cuobjdump -sass, no GPU needed): the mnemonics that prove the
Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, UTMALDG = TMA load, LDTM / STTM = tcgen05.ld / st,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA would be the legacy mma.sync path.
  python tools/sass_summary.py > profiles/r2_sass_opcodes.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "opencv-simpleslam_b200", "libb200slam.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "MUFU.EX2", "FFMA"]
rows, cur, cnt, total = [], None, None, collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    if cur:
        for op in OPS:
            if re.search(r"\b" + re.escape(op) + r"\b|\b" + re.escape(op) + r"\.", line):
                cnt[op] += 1
if cur:
    rows.append((cur, cnt))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode counts per kernel (cuobjdump -sass opencv-simpleslam_b200/libb200slam.so, sm_100a)\n")
print("| kernel | " + " | ".join(OPS) + " |")
print("|---|" + "---|" * len(OPS))
for (raw, c), name in sorted(zip(rows, names), key=lambda t: -t[0][1]["UTCHMMA"]):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("b2s::", "")
    if not any(c[o] for o in OPS[:10]) and "k_" not in name:
        continue
    total.update(c)
    print(f"| `{name}` | " + " | ".join(str(c[o]) for o in OPS) + " |")
print("| **all kernels** | " + " | ".join(str(total[o]) for o in OPS) + " |")
