"""Round-2 timeline of CTA 0 of the tcgen05 GEMM in its TMA-fed mode (MODE 3) on the conformer shapes (trace build:
APS_B200_VARIANT=trace APS_B200_NVCC_EXTRA=-DAPSB_TC_TRACE python -m aps_b200.build; APS_B200_LIB=aps_b200/libaps_b200_trace.so).
Per tile: when the MMAs of its k-blocks started, when the accumulator was committed, when the epilogue started / released
TMEM / finished — cycles from kernel start."""
import ctypes
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import _lib, ops  # noqa: E402

NAMES = {1: "tma_issue", 2: "mma_start", 3: "tile_commit", 4: "epi_start", 5: "epi_release", 6: "epi_done", 9: "start"}
dev = "cuda:0"
lib = _lib.load()
lib.aps_b200_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
M = 3200
CASES = [("ffn_a swish +lo", 256, 2048, "swish", 1, "128", True), ("ffn_a swish", 256, 2048, "swish", 1, "256", False),
         ("qkv", 256, 768, "none", 1, "64", False), ("out_proj +res", 256, 256, "none", 1, "64", False),
         ("ffn_b k5", 2048, 256, "none", 5, "128", False), ("ffn_b k5", 2048, 256, "none", 5, "256", False)]
for name, K, N, act, ks, bn, lo in CASES:
    os.environ["APS_B200_TC_BN"] = bn
    x = th.randn(M, K, device=dev)
    xl = ops.lo_companion(x)
    w, b = th.randn(N, K, device=dev) / K**0.5, th.randn(N, device=dev)
    cache = ops.SplitCache()
    if ks > 1:
        fn = lambda: ops.linear(x, w, None, cache=cache, x_lo=xl, ksplit=ks)
    else:
        fn = lambda: ops.linear(x, w, b, act=act, cache=cache, x_lo=xl, want_lo=lo)
    for _ in range(3):
        fn()
    buf = th.zeros(1024, dtype=th.int64, device=dev)
    lib.aps_b200_tc_trace(buf.data_ptr(), 1024)
    fn()
    th.cuda.synchronize()
    lib.aps_b200_tc_trace(None, 0)
    h = buf.cpu().tolist()
    n = min(h[0] & 0xFFFFFFFF, 1015)
    ev = sorted(((v & ((1 << 48) - 1)), (v >> 48) & 0xFFFF) for v in h[1:1 + n])
    t0 = ev[0][0]
    print(f"\n=== {name}: M={M} K={K} N={N} BN={bn} ksplit={ks}: {n} events, CTA 0 busy for {ev[-1][0] - t0} clk")
    by = {}
    for t, e in ev:
        by.setdefault(e, []).append(t - t0)
    for e in (1, 2, 3, 4, 5, 6):
        ts = by.get(e, [])
        if not ts:
            continue
        d = sorted(b_ - a_ for a_, b_ in zip(ts, ts[1:])) or [0]
        print(f"  {NAMES[e]:>12}: {len(ts):3d} events, first {ts[0]:6d}, last {ts[-1]:6d}, median spacing {d[len(d) // 2]:5d}")
    for i, (c, s4, s5, s6) in enumerate(zip(by.get(3, []), by.get(4, []), by.get(5, []), by.get(6, []))):
        print(f"  tile {i}: accumulator committed {c:6d} | epilogue start {s4:6d}  tmem released {s5:6d}  done {s6:6d}  (epilogue {s6 - s4} clk)")
