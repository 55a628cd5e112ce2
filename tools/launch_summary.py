"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [out.json]"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
head = rows[hi]
ki, mi, vi, ui, ii = (head.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    d = per.setdefault(r[ii], {'name': re.sub(r'\(.*', '', r[ki]).replace('void <unnamed>::', '').replace('<unnamed>::', '')})
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if r[mi].startswith('gpu__time'):
        d['us'] = v / 1e3 if u == 'ns' else (v if u == 'us' else v * 1e3)
    else:
        d[r[mi]] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d['name'], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get('us', 0)
    a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values())
out = []
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append({"kernel": k, "launches": a[0], "total_ms": a[1] / 1e3, "share": a[1] / tot, "avg_us": a[1] / a[0],
                "dram_mb_per_launch": a[2] / a[0] / 1e6, "dram_gbs": a[2] / max(a[1], 1e-9) / 1e3})
    print(f"{a[1] / 1e3:8.2f} ms {100 * a[1] / tot:5.1f}%  n={a[0]:4d}  avg {a[1] / a[0]:8.1f} us  dram {a[2] / a[0] / 1e6:8.1f} MB/launch "
          f"{a[2] / max(a[1], 1e-9) / 1e3:7.1f} GB/s  {k[:70]}")
print('total', tot / 1e3, 'ms')
if len(sys.argv) > 2:
    json.dump({"source": sys.argv[1], "total_ms": tot / 1e3, "note": "ncu per-launch times are cold-cache and serialised: compare shares",
               "kernels": out}, open(sys.argv[2], 'w'), indent=1)
