#!/usr/bin/env python
"""Static SASS evidence for profiles/: counts of the Blackwell-specific mnemonics per object file of libaps_b200.so
(`cuobjdump -sass build/*.o`): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor loads, UBLKCP = 1-D
bulk TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2 / FADD2 / FMUL2 = packed fp32, HMMA = legacy mma.sync (must be 0).

    python scripts/sass_listing.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "ACQBULK",
        "UCGABAR", "FFMA2", "FADD2", "FMUL2", "HMMA", "LDGSTS", "MUFU", "REDUX", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL"]


def main():
    objs = sorted(f for f in os.listdir(os.path.join(ROOT, "build")) if f.endswith(".o"))
    print("# cuobjdump -sass of build/*.o (sm_100a), static instruction counts; kernels per object in brackets")
    print(f"{'object':14s} {'kernels':>7s} {'instr':>8s}  " + " ".join(f"{k:>8s}" for k in KEYS))
    for o in objs:
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "build", o)], capture_output=True, text=True).stdout
        c = collections.Counter()
        kernels = out.count("Function : ")
        n = 0
        for line in out.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            n += 1
            op = m.group(1).split(".")[0]
            c[op] += 1
        print(f"{o:14s} {kernels:7d} {n:8d}  " + " ".join(f"{c.get(k, 0):8d}" for k in KEYS))
    # the tensor-core kernels one by one
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "build", "tc_gemm.o")], capture_output=True, text=True).stdout
    print("\n# tc_gemm.o per kernel: template <BN, MODE (0 linear gather, 1 conv2d gather, 2 conv_transpose2d, 3 TMA-fed linear, 4 TMA-fed conv2d), CL>")
    name, c = None, collections.Counter()
    rows = []
    for line in out.splitlines():
        if "Function : " in line:
            if name:
                rows.append((name, c))
            name, c = line.split("Function : ")[1].strip(), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            c[m.group(1).split(".")[0]] += 1
    if name:
        rows.append((name, c))
    for name, c in rows:
        k = re.search(r"tc_gemm_kernelILi(\d+)ELi(\d)ELi(\d)E", name)
        tag = f"tc_gemm_kernel<{k.group(1)},{k.group(2)},{k.group(3)}>" if k else name[:40]
        print(f"{tag:28s} instr {sum(c.values()):6d}  UTCHMMA {c.get('UTCHMMA', 0):3d}  UTCBAR {c.get('UTCBAR', 0):3d}  LDTM {c.get('LDTM', 0):3d}  "
              f"UTMALDG {c.get('UTMALDG', 0):3d}  SYNCS {c.get('SYNCS', 0):3d}  HMMA {c.get('HMMA', 0)}")


if __name__ == "__main__":
    main()
