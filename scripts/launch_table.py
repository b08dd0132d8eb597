#!/usr/bin/env python
"""Per-launch table of ONE bench step from an `ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...
--csv` launch list (times are cold-cache and serialised: compare SHARES).  A step = the launches between two consecutive
frontend_kernel launches that are more than 50 launches apart.

    python scripts/launch_table.py gpurun_out/r02c_launches_asr_encoder.csv [step index, default 5] > profiles/r02_....txt
"""
import collections
import csv
import re
import sys

TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
DUR = "gpu__time_duration.sum"


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    marker = sys.argv[3] if len(sys.argv) > 3 else "frontend_kernel"
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    idx = {h: i for i, h in enumerate(hdr)}
    data = collections.OrderedDict()
    for r in rd:
        if len(r) < len(hdr):
            continue
        d = data.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]], "grid": r[idx["Grid Size"]]})
        d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    items = list(data.values())
    fr = [i for i, d in enumerate(items) if marker in d["name"]]
    segs = [(fr[i], fr[i + 1]) for i in range(len(fr) - 1) if fr[i + 1] - fr[i] > 50]
    a, b = segs[min(which, len(segs) - 1)]
    step = items[a:b]
    tot = sum(d[DUR] for d in step)
    print(f"# {path}: step {which} = launches {a}..{b - 1} ({len(step)} launches, {tot / 1e3:.1f} us serialised under ncu)")
    short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("apsb::", "")[:52]
    print("\n## launches in order (first 40)")
    for i, d in enumerate(step[:40]):
        print(f"{i:3d} {short(d['name']):52s} grid {d['grid']:>14s} {d[DUR] / 1e3:8.1f} us  tensor pipe {d.get(TP, 0):5.1f} %")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in step:
        k = short(d["name"])
        agg[k][0] += 1
        agg[k][1] += d[DUR] / 1e3
        agg[k][2] += d.get(TP, 0) * d[DUR] / 1e3
    print("\n## by kernel: launches, us, share of the step, tensor-pipe active (time weighted)")
    ttot = 0.0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:52s} x{v[0]:3d} {v[1]:8.1f} us {100 * v[1] / (tot / 1e3):5.1f} %   tensor {v[2] / max(v[1], 1e-9):5.1f} %")
        ttot += v[2]
    print(f"\ntensor pipe active, time weighted over the whole step: {ttot / (tot / 1e3):.1f} %")


if __name__ == "__main__":
    main()
