#!/usr/bin/env python
"""Dynamic SASS histogram from `ncu -i X.ncu-rep --page source --csv --print-source sass` output."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
iW = hdr.index('L1 Wavefronts Shared'); iWi = hdr.index('L1 Wavefronts Shared Ideal')
c = collections.Counter(); s = collections.Counter(); w = collections.Counter(); wi = collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[iE].isdigit():
        continue
    src = r[iS].strip().split()
    op = src[0] if not src[0].startswith('@') else src[1]
    full = op
    op = op.split('.')[0]
    if op in ('LDS', 'STS', 'LDG', 'STG'):
        op = '.'.join(p for p in full.split('.') if p in (op, '64', '128', 'U8'))
    n = int(r[iE]); c[op] += n; tot += n; s[op] += int(r[iSm] or 0)
    w[op] += int(r[iW] or 0); wi[op] += int(r[iWi] or 0)
print('total warp inst', tot, 'per unit', tot / units, 'samples', sum(s.values()))
for op, n in c.most_common(40):
    print(f'{op:12s} {n:10d} {n/units:8.1f}/unit  samples {s[op]:6d}  smem wavefronts {w[op]} (ideal {wi[op]})')
