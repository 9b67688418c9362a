"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms "
      "(ncu: serialised, cold cache; compare SHARES)")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v:10.3f} ms {100 * v / tot:5.1f}%  x{n:<4d} {k}")
