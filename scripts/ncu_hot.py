"""Top stall-sample instructions of one kernel from `ncu -i rep --page source --csv` output.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python scripts/ncu_hot.py src.csv <substring> [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
b = [b for b in blocks if pat in b['name']][0]
hdr = b['rows'][0]
si = hdr.index('Source'); st = hdr.index('Warp Stall Sampling (All Samples)'); ie = hdr.index('Instructions Executed')
blk = [r for r in b['rows'][1:] if len(r) > max(si, st, ie)]
tot = sum(int(r[st] or 0) for r in blk); toti = sum(int(r[ie] or 0) for r in blk)
print(b['name'][:80]); print('n instr', len(blk), 'samples', tot, 'warp instr', toti)
top = sorted(range(len(blk)), key=lambda i: -int(blk[i][st] or 0))[:N]
for i in sorted(top):
    r = blk[i]
    print(f"{i:5d} {int(r[st]):6d} ({100*int(r[st])/max(tot,1):4.1f}%) exec={r[ie]:>9s}  {r[si].strip()[:100]}")
