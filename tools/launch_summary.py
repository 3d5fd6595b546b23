#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*] --csv): per-kernel totals of
the LAST step in the log (kernels of one bench step = launches / steps)."""
import csv, collections, sys
path = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    d.setdefault((int(r[ii]), r[ki].split('(')[0].replace('void ', '').replace('sn::', '')), {})[r[mi]] = float(r[vi].replace(',', ''))
seen = collections.OrderedDict()
for (i, k), m in d.items(): seen.setdefault(k, []).append(m)
out = []
for k, ms in seen.items():
    last = ms[-max(1, len(ms) // steps):]
    out.append((k, len(last), sum(x['gpu__time_duration.sum'] for x in last) / 1e6, sum(x.get('dram__bytes_read.sum', 0) for x in last) / 1e9, sum(x.get('dram__bytes_write.sum', 0) for x in last) / 1e9))
tot = sum(o[2] for o in out)
print(f"{'kernel':44s} {'launches':>8s} {'total_ms':>9s} {'share':>6s} {'dram_rd_GB':>10s} {'dram_wr_GB':>10s}")
for k, n, t, rb, wb in sorted(out, key=lambda o: -o[2]):
    print(f"{k[:44]:44s} {n:8d} {t:9.3f} {100*t/tot:5.1f}% {rb:10.2f} {wb:10.2f}")
print(f"{'TOTAL (kernels only)':44s} {sum(o[1] for o in out):8d} {tot:9.3f}")
