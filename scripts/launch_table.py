"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST
repetition (python scripts/launch_table.py file.csv [reps])."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
per = len(seq) // reps
agg = collections.OrderedDict()
for k, v in seq[-per:]:
    agg.setdefault(k[:64], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print('%-66s n=%d  %9.1f us  (%4.1f%%)' % (k, len(v), sum(v) / 1e3, 100 * sum(v) / tot))
print('total %.1f us over %d launches' % (tot / 1e3, per))
