"""A/B of the conformer-front convolutions (gather-fed tcgen05 GEMM) between library builds: APS_B200_LIB selects one."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402

dev = "cuda:0"
th.manual_seed(0)
cache = ops.SplitCache()
xc2 = th.randn(64, 199, 40, 256, device=dev)
wc2 = th.randn(256, 3, 3, 256, device=dev) / 48
xc3 = th.randn(64, 100, 20, 256, device=dev)
bc = th.randn(256, device=dev)


def timeit(fn, reps=4):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    g = th.cuda.CUDAGraph()
    with th.cuda.graph(g):
        for _ in range(reps):
            fn()
    best = 1e30
    for _ in range(5):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        th.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


t2 = timeit(lambda: ops.conv2d_nhwc(xc2, wc2, bc, stride=(2, 2), padding=(1, 1), act="relu", cache=cache))
t3 = timeit(lambda: ops.conv2d_nhwc(xc3, wc2, bc, stride=(2, 2), padding=(1, 1), act="relu", cache=cache))
print(f"{os.environ.get('APS_B200_LIB', 'default')}: conv2 {t2:8.1f} us   conv3 {t3:8.1f} us")
