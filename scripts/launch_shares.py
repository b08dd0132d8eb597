#!/usr/bin/env python
"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.Counter()
cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name)[:80]
    tot[name] += float(r[-1])
    cnt[name] += 1
s = sum(tot.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {s/1e3:.1f} us total (cold-cache, serialised: compare SHARES)")
for k, v in tot.most_common(14):
    print(f"{100*v/s:6.2f}%  {v/1e3:10.1f} us  x{cnt[k]:4d}  {k}")
