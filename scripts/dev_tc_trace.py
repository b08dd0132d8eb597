"""Timeline of CTA 0 of the persistent tcgen05 GEMM (needs the trace build:
APS_B200_VARIANT=trace APS_B200_NVCC_EXTRA=-DAPSB_TC_TRACE python -m aps_b200.build;  run with
APS_B200_LIB=aps_b200/libaps_b200_trace.so).  Prints per-event cycle stamps relative to kernel start."""
import ctypes
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import _lib, ops  # noqa: E402

NAMES = {1: "tma_issue", 2: "mma_start", 3: "mma_tile_commit", 4: "epi_start", 5: "epi_tmem_release", 6: "epi_done",
         7: "prod_stored", 8: "prod_slot_free", 9: "kernel_start", 11: "mma_a_ready"}
dev = "cuda:0"
lib = _lib.load()
lib.aps_b200_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
for (M, K, N, bn) in ((3200, 256, 2048, "256"), (3200, 256, 2048, "128"), (3200, 256, 256, "64"), (40000, 2304, 256, "256")):
    os.environ["APS_B200_TC_BN"] = bn
    x, w, b = th.randn(M, K, device=dev), th.randn(N, K, device=dev) / K**0.5, th.randn(N, device=dev)
    r = th.randn(M, N, device=dev)
    cache = ops.SplitCache()
    for _ in range(3):
        ops.linear(x, w, b, residual=r, cache=cache)
    buf = th.zeros(1024, dtype=th.int64, device=dev)
    lib.aps_b200_tc_trace(buf.data_ptr(), 1024)
    ops.linear(x, w, b, residual=r, cache=cache)
    th.cuda.synchronize()
    lib.aps_b200_tc_trace(None, 0)
    h = buf.cpu().tolist()
    n = min(h[0] & 0xFFFFFFFF, 1022)
    ev = sorted(((v & ((1 << 48) - 1)), (v >> 48) & 0xFFFF) for v in h[1:1 + n])
    t0 = ev[0][0]
    print(f"=== M={M} K={K} N={N} BN={bn}: {n} events, total {ev[-1][0] - t0} clk")
    last = {}
    out = []
    for t, e in ev[:160]:
        out.append(f"{t - t0:>8} {NAMES.get(e, e)}")
    print("\n".join(out))
    # per-event mean spacing
    for e in sorted(NAMES):
        ts = [t for t, k in ev if k == e]
        if len(ts) > 2:
            d = [b_ - a_ for a_, b_ in zip(ts, ts[1:])]
            print(f"  {NAMES[e]:>18}: {len(ts)} events, median spacing {sorted(d)[len(d)//2]} clk")
