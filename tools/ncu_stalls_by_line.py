#!/usr/bin/env python
"""Per-source-line roll-up of an ncu source page: stall samples (by reason) and executed instructions.
    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass -k regex:<kernel> --launch-count 1 > src.csv
    python tools/ncu_stalls_by_line.py src.csv [top_n]
(inlined frames are listed at every level, so a line's share includes the lines inlined into it)"""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur = None; hdr = None
agg = []; tot = 0; totinst = 0
f = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; H = len(hdr); continue
    if r[0] == '': continue   # sass row
    def col(name):
        i = hdr.index(name) - H
        return r[i]
    s = f(col('# Samples')); ins = f(col('Instructions Executed'))
    if s or ins:
        st = {}
        for i, k in enumerate(hdr):
            if k.startswith('stall_') and 'Not Issued' not in k:
                v = f(r[i - H])
                if v: st[k] = v
        agg.append((s, ins, cur, r[0], ' '.join(r[1:len(r) - H + 2])[:110], st))
        tot += s; totinst += ins
print('total samples', tot, 'inst', totinst)
allst = {}
for a in agg:
    for k, v in a[5].items(): allst[k] = allst.get(k, 0) + v
print(sorted(allst.items(), key=lambda x: -x[1]))
for a in sorted(agg, key=lambda x: -x[0])[:topn]:
    print(f"{a[0]*100/tot:5.1f}% inst {a[1]*100/totinst:5.1f}% {a[2]}:{a[3]} {a[4]}\n       {sorted(a[5].items(), key=lambda x:-x[1])[:4]}")
