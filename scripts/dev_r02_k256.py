"""Would the K = 256, N = 256 layers (attention out-projection, second point-wise conv) gain from two K slices + the
LayerNorm reduce, like the d_ff -> d_model layers do?  Device time from CUDA-graph replays."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402

dev = th.device("cuda", 0)
th.manual_seed(0)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    g = th.cuda.CUDAGraph()
    with th.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    th.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        th.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


M, K, N = 3200, 256, 256
cache = ops.SplitCache()
x = th.randn(M, K, device=dev)
xl = ops.lo_companion(x)
w = th.randn(N, K, device=dev) / 16
b = th.randn(N, device=dev)
res = th.randn(M, N, device=dev)
g, be = th.ones(N, device=dev), th.zeros(N, device=dev)


def plain():
    y = ops.linear(x, w, b, residual=res, cache=cache, x_lo=xl)
    return ops.layernorm2(y, g, be, 1e-5)


for bn in ("", "64", "128", "256"):
    if bn:
        os.environ["APS_B200_TC_BN"] = bn
    else:
        os.environ.pop("APS_B200_TC_BN", None)
    t_plain = timeit(plain)
    row = f"BN={bn or 'auto':>4s}: linear+res -> LN {t_plain:6.1f} us"
    for ks in (2, 4):
        def split():
            parts = ops.linear(x, w, None, cache=cache, x_lo=xl, ksplit=ks)
            return ops.layernorm2(parts, g, be, 1e-5, bias=b, residual=res)
        row += f" | k{ks} + LN-reduce {timeit(split):6.1f} us"
    print(row)
