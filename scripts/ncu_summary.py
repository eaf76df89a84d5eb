"""Summarise .ncu-rep captures (ncu --set full) into one text table: per launch duration, DRAM bytes / throughput,
tensor-pipe activity, registers, occupancy.  usage: python scripts/ncu_summary.py file.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active": "tensor_tf32_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_bytes.sum": "l2_bytes",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6,
        "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"== {path}")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        rec = {}
        for m, short in WANT.items():
            if m in col:
                v = r[col[m]].replace(",", "")
                try:
                    v = float(v) * UNIT.get(units[col[m]], 1)
                except ValueError:
                    pass
                rec[short] = v
        dur = rec.get("duration", 0.0)
        traffic = rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
        gbs = traffic / dur / 1e3 if dur else 0.0
        print(f"{name[:48]:48s} dur {dur:9.1f} us  dram r/w {rec.get('dram_read', 0) / 1e6:8.1f}/{rec.get('dram_write', 0) / 1e6:8.1f} MB"
              f" = {gbs:7.1f} GB/s ({rec.get('dram_pct', 0):5.1f}% peak)  tensor {rec.get('tensor_pct', 0)}%  sm {rec.get('sm_pct', 0)}%"
              f"  occ {rec.get('occupancy_pct', 0)}%  regs {rec.get('regs', 0)}  grid {rec.get('grid', 0)} x {rec.get('block', 0)}"
              f"  L2 {rec.get('l2_bytes', 0) / 1e6:8.1f} MB")
