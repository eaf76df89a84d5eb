"""Per-launch durations of the grouped GEMM within one update, from an ncu launch list (gpu__time_duration.sum --csv):
prints the LAST complete update's GEMM launches in order.  usage: gemm_launch_times.py launches.csv [launches_per_update]"""
import csv
import sys

path = sys.argv[1]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 18
lines = [l for l in open(path) if not l.startswith("==")]
t = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum" or "gemm_tf32_grouped" not in row["Kernel Name"]:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    t.append(v)
last = t[-per:]
print(" ".join(f"{x:.1f}" for x in last), "| sum", round(sum(last), 1), "us")
