"""One conformer-sized TMA-fed GEMM (FFN-a: [3200 x 256] x [256 x 2048], swish, lo companion) for an ncu capture."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402

dev = "cuda:0"
M, K, N = 3200, 256, 2048
x = th.randn(M, K, device=dev)
xl = ops.lo_companion(x)
w, b = th.randn(N, K, device=dev) / 16, th.randn(N, device=dev)
cache = ops.SplitCache()
for _ in range(6):
    y = ops.linear(x, w, b, act="swish", cache=cache, x_lo=xl, want_lo=True)
th.cuda.synchronize()
print(float(y[0].abs().sum()))
