"""Cycle-accurate ablation (trace build): cycles per k-block of CTA 0 and wall time of the same launch, for the
debug modes (bit 0: no A stores, bit 1: no TMA loads) -> separates clock throttling from pipeline stalls."""
import ctypes
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
lib.aps_b200_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
for (M, K, N, bn) in ((128000, 2304, 256, "256"), (128000, 2304, 256, "128")):
    os.environ["APS_B200_TC_BN"] = bn
    x, w, b = th.randn(M, K, device=dev), th.randn(N, K, device=dev) / K**0.5, th.randn(N, device=dev)
    cache = ops.SplitCache()
    for dbg in (16, 19, 31):
        os.environ["APS_B200_TC_DBG"] = str(dbg)
        for _ in range(3):
            ops.linear(x, w, b, cache=cache)
        buf = th.zeros(1024, dtype=th.int64, device=dev)
        lib.aps_b200_tc_trace(buf.data_ptr(), 1024)
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        ops.linear(x, w, b, cache=cache)
        e1.record()
        th.cuda.synchronize()
        lib.aps_b200_tc_trace(None, 0)
        h = buf.cpu().tolist()
        n = min(h[0] & 0xFFFFFFFF, 1022)
        ev = sorted(((v & ((1 << 48) - 1)), (v >> 48) & 0xFFFF) for v in h[1:1 + n])
        ts = [t for t, k in ev if k == 3]
        d = [b_ - a_ for a_, b_ in zip(ts, ts[1:])]
        nkb = K // (16 if bn == "256" else 32)
        print(f"M={M} BN={bn} dbg{dbg}: wall {e0.elapsed_time(e1)*1e3:.0f} us, cycles per tile {d} -> per k-block "
              f"{[round(v / nkb) for v in d]}; total cycles {ev[-1][0] - ev[0][0]} -> {(ev[-1][0] - ev[0][0]) / e0.elapsed_time(e1) / 1e3:.0f} MHz", flush=True)
