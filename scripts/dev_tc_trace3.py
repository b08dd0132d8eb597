"""Epilogue ablation (trace build): cycles from epi_start to epi_done per tile for small-K GEMMs."""
import ctypes
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
lib.aps_b200_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
for (M, K, N, bn) in ((3200, 256, 2048, "128"), (3200, 256, 256, "64")):
    os.environ["APS_B200_TC_BN"] = bn
    x, w, b = th.randn(M, K, device=dev), th.randn(N, K, device=dev) / K**0.5, th.randn(N, device=dev)
    r = th.randn(M, N, device=dev)
    cache = ops.SplitCache()
    for dbg, use_res in ((16, True), (16 + 64 + 128, True)):
        os.environ["APS_B200_TC_DBG"] = str(dbg)
        f = lambda: ops.linear(x, w, b, residual=r if use_res else None, act="swish", cache=cache)
        for _ in range(3):
            f()
        buf = th.zeros(1024, dtype=th.int64, device=dev)
        lib.aps_b200_tc_trace(buf.data_ptr(), 1024)
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        th.cuda.synchronize()
        lib.aps_b200_tc_trace(None, 0)
        h = buf.cpu().tolist()
        n = min(h[0] & 0xFFFFFFFF, 1022)
        ev = sorted(((v & ((1 << 48) - 1)), (v >> 48) & 0xFFFF) for v in h[1:1 + n])
        t0 = ev[0][0]
        s = " ".join(f"{k}@{t - t0}" for t, k in ev if k in (9, 3, 4, 5, 6))
        print(f"M={M} K={K} N={N} BN={bn} dbg{dbg} res={use_res}: wall {e0.elapsed_time(e1)*1e3:.0f} us | {s}", flush=True)
