"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
items = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    short = re.sub(r"void |tfmq::", "", re.sub(r"\(.*", "", row["Kernel Name"]))
    agg[short][0] += 1
    agg[short][1] += v
    tot += v
    items.append((short, v, row.get("Grid Size", "")))
print(f"# {path}: total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:64]:64s} n={n:4d} total={t:9.1f}us avg={t / n:8.1f}us share={100 * t / tot:5.1f}%")
if len(sys.argv) > 2:
    print()
    for s, v, g in sorted(items, key=lambda x: -x[1])[: int(sys.argv[2])]:
        print(f"{s[:50]:50s} {v:8.1f}us grid={g}")
