"""Round-2 visit O microbenchmarks (device time from CUDA-graph replays): the write roofline of the thin front convolution
(a plain fill of the same 521 MB), the 48 x 48 attention tile against the 64 x 64 one at the BASELINE encoder's L = 48,
and the split-K LayerNorm reduce."""
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


# ---- write roofline: fill of the conv1 output
y = th.empty(64, 198, 39, 256, device=dev)
mb = y.numel() * 4 / 1e6
t_fill = timeit(lambda: y.zero_(), 4)
src = th.randn_like(y)
t_copy = timeit(lambda: y.copy_(src), 4)
xin = th.randn(64, 398, 80, 1, device=dev)
w1 = th.randn(256, 3, 3, 1, device=dev)
b1 = th.randn(256, device=dev)
t_thin = timeit(lambda: ops.conv2d_nhwc(xin, w1, b1, stride=(2, 2), act="relu"), 4)
os.environ["APS_B200_THIN_UNR"] = "2"
t_thin2 = timeit(lambda: ops.conv2d_nhwc(xin, w1, b1, stride=(2, 2), act="relu"), 4)
y2 = ops.conv2d_nhwc(xin, w1, b1, stride=(2, 2), act="relu")
os.environ.pop("APS_B200_THIN_UNR")
y1 = ops.conv2d_nhwc(xin, w1, b1, stride=(2, 2), act="relu")
print(f"thin3x3 two positions per trip: {t_thin2:.1f} us, bit-identical: {bool(th.equal(y1, y2))}")
print(f"conv1 output {mb:.0f} MB: fill {t_fill:.1f} us ({mb / t_fill * 1e3:.0f} GB/s), copy {t_copy:.1f} us "
      f"({2 * mb / t_copy * 1e3:.0f} GB/s r+w), thin3x3 {t_thin:.1f} us ({mb / t_thin * 1e3:.0f} GB/s)")

# ---- attention at the BASELINE encoder size
N, L, H, E = 64, 48, 4, 256
qkv = th.randn(N * L, 3 * E, device=dev)
pos = th.randn(2 * L - 1, E // H, device=dev)
for mode, kw in ((0, {}), (1, dict(pos=pos))):
    t48 = timeit(lambda: ops.mhsa(qkv, N, L, H, mode=mode, want_lo=True, **kw))
    a = ops.mhsa(qkv, N, L, H, mode=mode, **kw)
    os.environ["APS_B200_MHSA_TILE64"] = "1"
    t64 = timeit(lambda: ops.mhsa(qkv, N, L, H, mode=mode, want_lo=True, **kw))
    b = ops.mhsa(qkv, N, L, H, mode=mode, **kw)
    os.environ.pop("APS_B200_MHSA_TILE64")
    print(f"attention N={N} L={L} H={H} mode {mode}: 48-tile {t48:.1f} us, 64-tile {t64:.1f} us, bit-identical: {bool(th.equal(a, b))}")

# ---- LayerNorm reduce of 5 split-K slices
M = 3072
parts = th.randn(5, M, 256, device=dev)
g, be, b = th.ones(256, device=dev), th.zeros(256, device=dev), th.randn(256, device=dev)
r2 = th.randn(M, 256, device=dev)
print(f"layernorm2 5 slices [3072, 256]: {timeit(lambda: ops.layernorm2(parts, g, be, 1e-5, bias=b, residual=r2, alpha=0.5)):.1f} us; "
      f"1 slice: {timeit(lambda: ops.layernorm2(parts[0], g, be, 1e-5, bias=b, residual=r2, alpha=0.5)):.1f} us")
