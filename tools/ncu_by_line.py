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


def parse_dis(path, kernel_sub):
    lines = open(path).read().splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.startswith('.text.') and kernel_sub in l:
            start = i
            break
    assert start is not None, "kernel not found in disassembly"
    out = []            # (innermost (file, line), outermost-in-kernels.cu line)
    cur = None
    stack = []
    for l in lines[start + 1:]:
        if l.startswith('.text.') or l.startswith('\t.section'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            f, ln, rest = m.group(1), int(m.group(2)), m.group(3)
            if 'inlined at' in rest:
                stack.append((f, ln))
            else:
                stack = [(f, ln)]
            cur = list(stack)
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
            inner = cur[0] if cur else ('?', 0)
            outer = 0
            if cur:
                for f, ln in cur:
                    if f.endswith('lsm_kernels.cu'):
                        outer = ln
            out.append((inner, outer, l.strip()[:100]))
    return out


def main():
    sass_csv, dis, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    dis_rows = parse_dis(dis, ksub)
    rows = list(csv.reader(open(sass_csv)))
    # find the block for this kernel (first occurrence)
    k0 = None
    for i, r in enumerate(rows):
        if r and r[0] == 'Kernel Name' and ksub.replace('ILi0E', '<(int)0>').split('<')[0] in r[1]:
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
