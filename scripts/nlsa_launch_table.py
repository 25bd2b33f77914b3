"""Kernel-time table of the LAST NLSA.analyse call in an ncu launch list (ncu --metrics gpu__time_duration.sum --csv of
scripts/nlsa_timing.py):   python scripts/nlsa_launch_table.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
last = max(i for i, (k, v) in enumerate(seq) if 'k_nlsa_cond' in k)
agg = collections.OrderedDict()
for k, v in seq[last:]:
    k = k.split('(')[0].replace('void ', '').replace('mem::', '')[:56]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print('%-58s n=%4d %9.1f us (%4.1f%%)' % (k, a[0], a[1] / 1e3, 100 * a[1] / tot))
print('total %.1f us of kernel time over %d launches' % (tot / 1e3, sum(a[0] for a in agg.values())))
