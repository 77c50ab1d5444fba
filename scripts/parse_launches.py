"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per step."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [(re.sub(r"<.*", "", r["Kernel Name"]).replace("void ", "")[:48], float(r["Metric Value"])) for r in rows]
idx = [i for i, (n, _) in enumerate(names) if "k_plan_finish" in n]          # last kernel of a plan
a, b = idx[-2] + 1, idx[-1] + 1                                             # one step: kernels + the next plan
agg = OrderedDict()
for n, t in names[a:b]:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':50s} {'n':>3s} {'us':>9s} {'share':>7s}")
for n, (c, t) in agg.items():
    print(f"{n:50s} {c:3d} {t / 1000:9.2f} {100 * t / tot:6.1f}%")
print(f"{'TOTAL (one step, serialised, cold cache)':50s} {sum(v[0] for v in agg.values()):3d} {tot / 1000:9.2f}")
