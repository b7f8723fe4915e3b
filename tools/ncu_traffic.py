#!/usr/bin/env python
"""dram bytes (read + write) per launch of every captured kernel of an ncu report -> JSON on stdout.
usage: python tools/ncu_traffic.py prof.ncu-rep"""
import csv, json, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
res = {}
for r in rows[2:]:
    k = r[col('Kernel Name')].split('(')[0].replace('void ', '').replace('lsm::', '')
    rd = float(r[col('dram__bytes_read.sum')]) * scale[units[col('dram__bytes_read.sum')]]
    wr = float(r[col('dram__bytes_write.sum')]) * scale[units[col('dram__bytes_write.sum')]]
    res.setdefault(k, []).append({'read': rd, 'write': wr, 'us': float(r[col('gpu__time_duration.sum')])})
print(json.dumps({k: {'launches': len(v), 'dram_read_bytes': sum(x['read'] for x in v) / len(v),
                      'dram_write_bytes': sum(x['write'] for x in v) / len(v),
                      'duration_us_under_ncu': sum(x['us'] for x in v) / len(v)} for k, v in res.items()}, indent=1))
