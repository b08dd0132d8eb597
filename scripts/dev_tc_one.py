"""One small-K GEMM with a swish + residual epilogue, a few calls (target of an ncu source-level capture)."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402

dev = "cuda:0"
M, K, N = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (3200, 256, 2048)))
x, w, b = th.randn(M, K, device=dev), th.randn(N, K, device=dev) / K**0.5, th.randn(N, device=dev)
r = th.randn(M, N, device=dev)
cache = ops.SplitCache()
for _ in range(6):
    y = ops.linear(x, w, b, residual=r, act="swish", cache=cache)
th.cuda.synchronize()
print(float(y.abs().mean()))
