"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list for profiles/.

    python tools/launch_summary.py gpurun_out/launches.csv "title" > profiles/<name>.txt
"""
import csv
import sys
from collections import OrderedDict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        rows.append((r["Kernel Name"], us))
tot = sum(u for _, u in rows)
agg = OrderedDict()
for k, u in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += u
print(f"# {title}")
print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
print(f"# launches captured: {len(rows)}, total {tot / 1e3:.3f} ms")
print("# kernel | launches | total us | share | mean us")
for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:100]} | {n} | {u:.1f} | {100 * u / tot:.1f}% | {u / n:.1f}")
print("# first 40 launches in order (kernel, us)")
for k, u in rows[:40]:
    print(f"{k[:80]}  {u:.1f}")
