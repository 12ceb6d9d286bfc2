"""Per-launch table of an `ncu --csv --metrics ...` log (duration, DRAM bytes, L2 hit rate, occupancy, instructions).
usage: python profiles/launch_table.py gpurun_out/x.csv"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); mi = H.index('Metric Name'); vi = H.index('Metric Value'); ii = H.index('ID')
d = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    d.setdefault(r[ii], {'k': r[ki][5:45]})[r[mi]] = r[vi].replace(',', '')
def g(v, m, sc=1.0):
    return float(v[m]) / sc if m in v else float('nan')
for k, v in d.items():
    print(k, v['k'], 'ms=%.3f' % g(v, 'gpu__time_duration.sum', 1e6), 'rdGB=%.2f' % g(v, 'dram__bytes_read.sum', 1e9), 'wrGB=%.2f' % g(v, 'dram__bytes_write.sum', 1e9),
          'hit=%.1f' % g(v, 'lts__t_sector_hit_rate.pct'), 'warps=%.1f' % g(v, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
          'Minst=%.0f' % g(v, 'smsp__inst_executed.sum', 1e6), 'issue=%.1f' % g(v, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
          'Msect=%.0f' % g(v, 'lts__t_sectors.sum', 1e6), 'thr/inst=%.1f' % g(v, 'smsp__thread_inst_executed_per_inst_executed.ratio'))
