"""One small-N transposed convolution of the DCCRN decoder (target of an ncu capture)."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402

dev = "cuda:0"
B, H, W, Ci, Co = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (32, 65, 251, 128, 32)))
x = th.randn(B, H, W, Ci, device=dev)
skip = th.randn(B, H, W, Ci, device=dev) if len(sys.argv) > 6 and sys.argv[6] == "skip" else None
w = th.randn(Co, 3, 3, Ci * (2 if skip is not None else 1), device=dev) * 0.05
b = th.randn(Co, device=dev)
cache = ops.SplitCache()
for _ in range(5):
    y = ops.conv_transpose2d_nhwc(x, w, b, stride=(2, 1), padding=(1, 1), output_padding=(0, 0), act="leaky_relu", leaky=0.01,
                                  cache=cache, skip=skip)
th.cuda.synchronize()
print(tuple(y.shape), float(y.abs().mean()))
