"""Summarise an ncu --csv launch list (per-kernel metrics) into a table."""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:28]), {})[r[im]] = r[iv].replace(',', '')
tot = 0.0
for k, v in d.items():
    ms = float(v['gpu__time_duration.sum']) / 1e6
    tot += ms
    g = lambda n, s=1.0: (float(v[n]) / s) if n in v else float('nan')
    print('%3s %-28s %7.3f ms regs %3s warps%% %5.1f dramGB %5.2f noinst %5.2f issue%% %5.1f' % (
        k[0], k[1], ms, v.get('launch__registers_per_thread', '?'), g('sm__warps_active.avg.pct_of_peak_sustained_active'),
        g('dram__bytes_read.sum', 1e9) + g('dram__bytes_write.sum', 1e9),
        g('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'), g('smsp__issue_active.avg.pct_of_peak_sustained_active')))
print('total ms', tot)
