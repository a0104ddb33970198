#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn, mu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        name = r[kn].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"launches: {sum(v[0] for v in agg.values())}, total device time {tot:.1f} us "
          "(ncu per-launch times are cold-cache and serialised: compare SHARES)\n")
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
