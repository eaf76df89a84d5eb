"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    if row["Metric Unit"] == "ns":
        v /= 1000
    elif row["Metric Unit"] == "ms":
        v *= 1000
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:10.1f} us {100*v/T:5.1f}%  n={cnt[k]:3d}  avg={v/cnt[k]:8.1f} us  {k}")
print(f"total {T:.1f} us over {sum(cnt.values())} launches")
