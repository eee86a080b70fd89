"""Summarise an ncu report (raw page CSV) per kernel launch:  python scripts/ncu_summary.py rep.ncu-rep > summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [c for c in WANT if c in idx]
print("kernel | " + " | ".join(f"{c} [{units[idx[c]]}]" for c in cols))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[idx["Kernel Name"]]
    print(name[:70] + " | " + " | ".join(r[idx[c]] for c in cols))
