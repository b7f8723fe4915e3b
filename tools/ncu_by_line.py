#!/usr/bin/env python
"""Correlate an ncu SASS-level source page with CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all liblsm_b200.so ; nvdisasm -g -c lsm_kernels.sm_100a.cubin > dis.txt
    python tools/ncu_by_line.py sass.csv dis.txt <mangled-kernel-substring> [regions.txt]

Joins by instruction order inside the kernel (ncu lists the SASS in program order, nvdisasm -g
annotates every instruction with `//## File "...", line N`, inlined frames included: the innermost
line is used, plus the outermost line inside lsm_kernels.cu for the per-phase roll-up).
"""
import csv
import re
import sys


OUTER_FILE = "lsm_kernels.cu"


def parse_dis(path, kernel_sub):
    """path: output of `nvdisasm -gi -c <cubin>` (inline chains) or `-g` (innermost only)."""
    lines = open(path).read().splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.startswith('.text.') and kernel_sub in l:
            start = i
            break
    assert start is not None, "kernel not found in disassembly"
    out = []            # (innermost (file, line), outermost line inside OUTER_FILE, sass text)
    cur = []            # annotation block attached to the next instruction(s)
    fresh = True
    pat = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    for l in lines[start + 1:]:
        if l.startswith('.text.') or l.startswith('\t.section'):
            break
        m = pat.search(l)
        if m:
            if fresh:
                cur = []
                fresh = False
            cur.append((m.group(1), int(m.group(2))))
            if m.group(3):
                cur.append((m.group(3), int(m.group(4))))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
            fresh = True
            inner = cur[0] if cur else ('?', 0)
            outer = 0
            for f, ln in cur:
                if f.endswith(OUTER_FILE):
                    outer = ln
            out.append((inner, outer, l.strip()[:100]))
    return out


def main():
    global OUTER_FILE
    sass_csv, dis, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    if len(sys.argv) > 5:
        OUTER_FILE = sys.argv[5]
    dis_rows = parse_dis(dis, ksub)
    rows = list(csv.reader(open(sass_csv)))
    # find the block for this kernel (first occurrence)
    k0 = None
    nsub = sys.argv[6] if len(sys.argv) > 6 else ksub.replace('ILi0E', '<(int)0>').split('<')[0]
    for i, r in enumerate(rows):
        if r and r[0] == 'Kernel Name' and nsub in r[1]:
            k0 = i
            break
    if k0 is None:
        k0 = 0
    hdr = rows[k0 + 1]
    ci, si, ti = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
    data = []
    for r in rows[k0 + 2:]:
        if not r or r[0] == 'Kernel Name':
            break
        data.append((float(r[ci] or 0), float(r[si] or 0), float(r[ti] or 0), r[1]))
    n = min(len(data), len(dis_rows))
    print(f"# sass rows ncu={len(data)} nvdisasm={len(dis_rows)}")
    by_outer = {}
    by_inner = {}
    tot_i = sum(d[0] for d in data[:n]); tot_s = sum(d[1] for d in data[:n])
    for k in range(n):
        (inner, outer, _), (ins, smp, tins, _) = dis_rows[k], data[k]
        a = by_outer.setdefault(outer, [0, 0, 0]); a[0] += ins; a[1] += smp; a[2] += tins
        key = (inner[0].split('/')[-1], inner[1])
        b = by_inner.setdefault(key, [0, 0, 0]); b[0] += ins; b[1] += smp; b[2] += tins
    print(f"# total warp-instructions {tot_i:.0f}, samples {tot_s:.0f}")
    regions = []
    if len(sys.argv) > 4:
        for l in open(sys.argv[4]):
            a, b, name = l.strip().split(None, 2)
            regions.append((int(a), int(b), name))
    if regions:
        print("\n## by phase (outermost line in lsm_kernels.cu)")
        for a, b, name in regions:
            i = sum(v[0] for k, v in by_outer.items() if a <= k <= b)
            s = sum(v[1] for k, v in by_outer.items() if a <= k <= b)
            t = sum(v[2] for k, v in by_outer.items() if a <= k <= b)
            print(f"{name:42s} inst {100 * i / tot_i:5.1f}%  stall-samples {100 * s / tot_s:5.1f}%  lanes/inst {t / max(i, 1):5.1f}")
    print("\n## top outer lines by samples")
    for k, v in sorted(by_outer.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"line {k:5d} inst {100 * v[0] / tot_i:5.2f}% samples {100 * v[1] / tot_s:5.2f}% lanes/inst {v[2] / max(v[0], 1):5.1f}")
    print("\n## top inner (file,line) by samples")
    for k, v in sorted(by_inner.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{k[0]}:{k[1]:<5d} inst {100 * v[0] / tot_i:5.2f}% samples {100 * v[1] / tot_s:5.2f}%")


if __name__ == '__main__':
    main()
