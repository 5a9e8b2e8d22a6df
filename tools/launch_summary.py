"""Aggregate an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) by kernel: launches, total ms, share,
DRAM GB.  usage: launch_summary.py launches.csv [first_id last_id]"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
per = OrderedDict()
for r in rows[1:]:
    i = int(r[iid])
    if i < lo or i > hi:
        continue
    per.setdefault((i, r[ik]), {})[r[im]] = float(r[iv].replace(',', ''))
agg = OrderedDict()
for (i, k), v in per.items():
    name = re.sub(r'\(.*$', '', k)
    name = re.sub(r'^void lmc::', '', name)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v.get('gpu__time_duration.sum', 0.0) / 1e6
    a[2] += (v.get('dram__bytes_read.sum', 0.0) + v.get('dram__bytes_write.sum', 0.0)) / 1e9
tot = sum(a[1] for a in agg.values())
totb = sum(a[2] for a in agg.values())
print('%-60s %6s %10s %7s %9s' % ('kernel', 'n', 'ms', 'share', 'dram GB'))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-60s %6d %10.3f %6.1f%% %9.3f' % (k[:60], a[0], a[1], 100 * a[1] / tot, a[2]))
print('%-60s %6d %10.3f %6.1f%% %9.3f' % ('TOTAL', sum(a[0] for a in agg.values()), tot, 100.0, totb))
