"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel: python tools/summarize_launches.py file.csv [skip]"""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg, tot, i = collections.OrderedDict(), 0.0, 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    i += 1
    if i <= skip:
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("b2s::", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    key = f"{name} grid{row.get('Grid Size', '')}"
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"# total {tot:.1f} us over {i - skip} launches")
print("kernel,launches,total_us,avg_us,share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{t:.1f},{t / n:.2f},{t / tot:.4f}")
