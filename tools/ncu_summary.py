"""Key metrics + stall breakdown of one ncu report (first kernel)."""
import csv, subprocess, sys
rep = sys.argv[1]
# a .ncu-rep report, or the csv of its raw page (`ncu -i rep --page raw --csv`, what tools/ncu_hot.sh brings back)
out = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
def g(k):
    try: return float(m[k].replace(',', ''))
    except Exception: return float('nan')
print('kernel', m.get('Kernel Name', '?')[:60])
unit_t = dict(zip(hdr, units)).get('gpu__time_duration.sum', 'ns')
t_ms = g('gpu__time_duration.sum') * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3, 'nsecond': 1e-6}.get(unit_t, 1e-6)
print('time ms %.3f  regs %s  warps_active%% %.1f  issue_active%% %.1f  thread_inst/inst %.1f' % (
    t_ms, m.get('launch__registers_per_thread'),
    g('sm__warps_active.avg.pct_of_peak_sustained_active'), g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
    g('smsp__thread_inst_executed_per_inst_executed.ratio')))
ud = dict(zip(hdr, units))
def gb(k):
    return g(k) * {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}.get(ud.get(k, 'byte'), 1e-9)
print('dram rd %.3f GB wr %.3f GB  dram%% %.1f  l1 hit %.1f  l2 hit %.1f  inst %.3g' % (gb('dram__bytes_read.sum'), gb('dram__bytes_write.sum'),
      g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), g('l1tex__t_sector_hit_rate.pct'), g('lts__t_sector_hit_rate.pct'), g('smsp__inst_executed.sum')))
stalls = {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''): g(k) for k in hdr if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')}
print('stalls per issue:', ', '.join('%s %.2f' % kv for kv in sorted(stalls.items(), key=lambda x: -x[1])[:8]))
for k in ['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_global_loads', 'sass__inst_executed_global_stores', 'smsp__inst_executed_op_branch.sum']:
    if k in m: print('  ', k, m[k])
