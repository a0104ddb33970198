"""Every launch of an `ncu --set full` raw page, in order, with the roofline-relevant columns.

    ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_rows.py raw.csv [label ...] > profiles/rNN_x.csv
"""
import csv, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
rows = list(csv.reader(open(sys.argv[1])))
labels = sys.argv[2:]
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
kn = hdr.index("Kernel Name")
cols = [c for c in COLS if c in hdr]
idx = [hdr.index(c) for c in cols]
w = csv.writer(sys.stdout)
w.writerow(["#", "label", "Kernel Name"] + cols)
w.writerow(["", "", ""] + [units[i] for i in idx])
for n, r in enumerate(data):
    if len(r) > max(idx):
        w.writerow([n, labels[n] if n < len(labels) else "", r[kn].replace("void ", "").replace("vae::", "")[:60]] +
                   [r[i].replace(",", "") for i in idx])
