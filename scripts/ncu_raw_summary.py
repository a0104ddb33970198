"""Per-kernel summary of an `ncu --set full` report (last launch of each kernel).

    ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_raw_summary.py raw.csv > profiles/rNN_ncu_full_top_kernels.csv
"""
import csv, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
kn = hdr.index("Kernel Name")
idx = [hdr.index(c) for c in COLS]
last = {}
for r in data:
    if len(r) > max(idx):
        last[r[kn].replace("void ", "").replace("vae::", "")] = r   # second round of the script = warmed-up launch
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + COLS)
w.writerow([""] + [units[i] for i in idx])
for k, r in last.items():
    w.writerow([k] + [r[i].replace(",", "") for i in idx])
