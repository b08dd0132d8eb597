"""Mainloop ablation on the trace build: time big GEMMs with the A stores and / or the TMA loads disabled."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402

dev = "cuda:0"


def timed(f, n=10):
    for _ in range(3):
        f()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (M, K, N) in ((128000, 2304, 256), (32000, 2304, 512)):
    x, w, b = th.randn(M, K, device=dev), th.randn(N, K, device=dev) / K**0.5, th.randn(N, device=dev)
    cache = ops.SplitCache()
    for bn in ("256", "128", "64"):
        os.environ["APS_B200_TC_BN"] = bn
        line = f"M={M} K={K} N={N} BN={bn}:"
        for dbg in (0, 1, 2, 3):
            os.environ["APS_B200_TC_DBG"] = str(dbg)
            t = timed(lambda: ops.linear(x, w, b, cache=cache))
            line += f"  dbg{dbg} {t*1e3:.0f} us"
        print(line, flush=True)
