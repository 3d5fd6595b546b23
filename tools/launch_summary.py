#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv).
usage: tools/launch_summary.py launches.csv [skip_first_n_launches]   (the warm-up step of bench.py --steps 1 --warmup 1 is skipped
by taking the LAST launch group: pass the number of launches of one step to keep only the tail)"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if r[0] == "ID")
data = [r for r in rows if r[0].isdigit()]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = collections.OrderedDict()
for r in data:
    per.setdefault(r[0], {"name": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
launches = list(per.values())
if len(sys.argv) > 2:
    launches = launches[-int(sys.argv[2]):]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for l in launches:
    name = re.sub(r"\(.*", "", l["name"]); name = re.sub(r"^void (sn::|snc::)?", "", name).replace("sn::", "")
    a = agg[name]
    a[0] += 1; a[1] += l.get("gpu__time_duration.sum", 0.0); a[2] += l.get("dram__bytes_read.sum", 0.0); a[3] += l.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("launches %d  total %.3f ms  dram read %.2f GB  write %.2f GB" % (len(launches), tot / 1e6, sum(a[2] for a in agg.values()) / 1e9, sum(a[3] for a in agg.values()) / 1e9))
print("%-44s %8s %9s %6s %10s %10s" % ("kernel", "launches", "total_ms", "share", "dram_rd_GB", "dram_wr_GB"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %8d %9.3f %5.1f%% %10.2f %10.2f" % (k[:44], a[0], a[1] / 1e6, 100 * a[1] / tot, a[2] / 1e9, a[3] / 1e9))
