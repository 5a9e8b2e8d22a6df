"""Aggregate ncu per-instruction samples (--page source --csv) by device function using the
cubin symbol table.  usage: sass_hot.py <src.csv> <cubin> <kernel-substring>"""
import csv, re, subprocess, sys
src, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, isamp, iinst, ithr = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
data = [(int(r[ia], 16), int(r[isamp]), int(r[iinst]), int(r[ithr])) for r in rows[2:] if len(r) > ithr]
base = data[0][0]
# symbols
out = subprocess.run(['readelf', '-sW', cubin], capture_output=True, text=True).stdout
secs = subprocess.run(['readelf', '-SW', cubin], capture_output=True, text=True).stdout
# find section index of kernel text
sec_idx = None
for line in secs.splitlines():
    m = re.match(r'\s*\[\s*(\d+)\]\s+(\S+)', line)
    if m and m.group(2).startswith('.text.') and kname in m.group(2):
        sec_idx = m.group(1)
syms = []
for line in out.splitlines():
    p = line.split()
    if len(p) >= 8 and p[3] == 'FUNC' and p[6] == sec_idx:
        syms.append((int(p[1], 16), int(p[2], 0), p[7]))
syms.sort()
def demangle(n):
    n = n.split('$')[-1] if '$' in n else n
    return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()[:70]
agg = {}
for addr, s, ins, thr in data:
    off = addr - base
    name = 'kernel-body'
    for o, sz, nm in syms:
        if o <= off < o + sz:
            name = nm
    a = agg.setdefault(name, [0, 0, 0])
    a[0] += s; a[1] += ins; a[2] += thr
tot = sum(a[0] for a in agg.values())
for nm, a in sorted(agg.items(), key=lambda x: -x[1][0])[:22]:
    print('%5.1f%% samples  inst %9d  lanes/inst %4.1f  %s' % (100.0 * a[0] / tot, a[1], a[2] / max(1, a[1]), demangle(nm)))
