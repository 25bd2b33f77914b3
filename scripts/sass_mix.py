"""Instruction mix of one kernel from `ncu --page source --csv` (python scripts/sass_mix.py rep.ncu-rep kernel_regex)."""
import collections
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '-k', 'regex:' + sys.argv[2]],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
si, ei = h.index('Source'), h.index('Instructions Executed')
wi = h.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in h else None
ops = collections.Counter()
wav = collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ei or r[0] == 'Address' or r[0] == 'Kernel Name':
        break
    try:
        v = float(r[ei])
    except ValueError:
        continue
    toks = r[si].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LD', 'ST', 'F2', 'I2', 'D')) and '.' in op else '')
    ops[op] += v
    tot += v
    if wi is not None:
        try:
            wav[op] += float(r[wi])
        except ValueError:
            pass
print('total warp-instructions %.3e' % tot)
for k, v in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 28):
    print('%-14s %12d  %5.1f%%   shared wavefronts %d' % (k, v, 100 * v / tot, wav.get(k, 0)))
