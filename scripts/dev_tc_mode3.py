"""Round-2 microbenchmark: conformer-sized linear layers on the tcgen05 engine, gather-fed (MODE 0) against TMA-fed
(MODE 3, raw x + lo companion), every tile width, and split-K for the skinny shapes.  Prints us per launch and the
fp32-equivalent TFLOP/s; numbers go to profiles/r02_tc_mode3_microbench.txt."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402

dev = th.device("cuda", 0)
th.manual_seed(0)
M = 3200
SHAPES = [("ffn_a swish", 256, 2048, "swish", 1), ("ffn_b", 2048, 256, "none", 1), ("ffn_b k4", 2048, 256, "none", 4),
          ("ffn_b k5", 2048, 256, "none", 5), ("ffn_b k8", 2048, 256, "none", 8), ("qkv", 256, 768, "none", 1),
          ("out_proj", 256, 256, "none", 1), ("pw1 glu", 256, 512, "glu", 1), ("front", 2560, 256, "none", 1),
          ("front k5", 2560, 256, "none", 5)]


def timeit(fn, reps=20):
    """us per call on the DEVICE: `reps` calls captured into one CUDA graph (the Python / ctypes cost of a call, ~10 us,
    would otherwise hide kernels shorter than that), replayed 5 times, best replay."""
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


cache = ops.SplitCache()
print(f"{'shape':14s} {'K':>5s} {'N':>5s} {'BN':>4s} {'mode':>6s} {'us':>8s} {'TF/s fp32-eq':>12s}")
for name, K, N, act, ks in SHAPES:
    x = th.randn(M, K, device=dev)
    xl = ops.lo_companion(x)
    w = th.randn(N, K, device=dev) / K**0.5
    b = th.randn(N, device=dev)
    res = th.randn(M, N, device=dev) if act == "none" else None
    fl = 2.0 * M * K * N
    for bn in ("64", "128", "256"):
        if int(bn) > 64 and N <= int(bn) // 2:
            continue
        os.environ["APS_B200_TC_BN"] = bn
        if ks == 1:
            t0 = timeit(lambda: ops.linear(x, w, b, act=act, residual=res, cache=cache))
            t3 = timeit(lambda: ops.linear(x, w, b, act=act, residual=res, cache=cache, x_lo=xl))
            t3l = timeit(lambda: ops.linear(x, w, b, act=act, residual=res, cache=cache, x_lo=xl, want_lo=True)) if act != "glu" else float("nan")
            print(f"{name:14s} {K:5d} {N:5d} {bn:>4s} {'gather':>6s} {t0:8.1f} {fl / t0 / 1e6:12.1f}")
            print(f"{name:14s} {K:5d} {N:5d} {bn:>4s} {'tma':>6s} {t3:8.1f} {fl / t3 / 1e6:12.1f}")
            print(f"{name:14s} {K:5d} {N:5d} {bn:>4s} {'tma+lo':>6s} {t3l:8.1f} {fl / t3l / 1e6:12.1f}")
        else:
            g, be = th.ones(N, device=dev), th.zeros(N, device=dev)
            r2 = th.randn(M, N, device=dev)
            tk = timeit(lambda: ops.linear(x, w, None, cache=cache, x_lo=xl, ksplit=ks))
            parts = ops.linear(x, w, None, cache=cache, x_lo=xl, ksplit=ks)
            tl = timeit(lambda: ops.layernorm2(parts, g, be, 1e-5, bias=b, residual=r2, alpha=0.5))
            print(f"{name:14s} {K:5d} {N:5d} {bn:>4s} {'splitk':>6s} {tk:8.1f} {fl / tk / 1e6:12.1f}   + layernorm2 reduce {tl:6.1f} us")
os.environ.pop("APS_B200_TC_BN", None)
# plain LayerNorm kernels for comparison
x = th.randn(M, 256, device=dev)
r = th.randn(M, 256, device=dev)
g, be = th.ones(256, device=dev), th.zeros(256, device=dev)
print(f"layernorm (old)  {timeit(lambda: ops.layernorm(x, g, be, 1e-5, residual=r, alpha=0.5)):6.1f} us;  layernorm2 (1 part, +lo) "
      f"{timeit(lambda: ops.layernorm2(x, g, be, 1e-5, residual=r, alpha=0.5)):6.1f} us")
# the thin front convolution
xin = th.randn(64, 398, 80, 1, device=dev)
w1 = th.randn(256, 3, 3, 1, device=dev)
b1 = th.randn(256, device=dev)
t_new = timeit(lambda: ops.conv2d_nhwc(xin, w1, b1, stride=(2, 2), padding=(1, 1), act="relu"), 4)
os.environ["APS_B200_NO_THIN_CONV"] = "1"
t_old = timeit(lambda: ops.conv2d_nhwc(xin, w1, b1, stride=(2, 2), padding=(1, 1), act="relu"), 4)
os.environ.pop("APS_B200_NO_THIN_CONV")
print(f"front conv1 [64,398,80,1] -> 256 ch: thin3x3 {t_new:7.1f} us, conv2d_narrow {t_old:7.1f} us (521 MB written: {521.6 / t_new * 1e3:6.0f} GB/s)")

# ---- weight-tile multicast across a cluster of CL CTAs (round 2): the key shapes at the tile width the cost model picks
os.environ.pop("APS_B200_TC_BN", None)
print("\ncluster size sweep (APS_B200_TC_CL), us per launch")
print(f"{'shape':22s} {'CL=1':>8s} {'CL=2':>8s} {'CL=4':>8s}")
xc2 = th.randn(64, 199, 40, 256, device=dev)
wc2 = th.randn(256, 3, 3, 256, device=dev) / 48
xc3 = th.randn(64, 100, 20, 256, device=dev)
bc = th.randn(256, device=dev)
rows = []
for name, K, N, act, ks in SHAPES:
    if "k4" in name or "k8" in name or (name in ("ffn_b", "front")):
        continue
    x = th.randn(M, K, device=dev)
    xl = ops.lo_companion(x)
    w = th.randn(N, K, device=dev) / K**0.5
    b = th.randn(N, device=dev)
    if ks == 1:
        rows.append((name + " tma", lambda x=x, w=w, b=b, act=act, xl=xl: ops.linear(x, w, b, act=act, cache=cache, x_lo=xl)))
        rows.append((name + " gather", lambda x=x, w=w, b=b, act=act: ops.linear(x, w, b, act=act, cache=cache)))
    else:
        rows.append((name, lambda x=x, w=w, xl=xl, ks=ks: ops.linear(x, w, None, cache=cache, x_lo=xl, ksplit=ks)))
rows.append(("conv2 3x3 s2 [64,199,40,256]", lambda: ops.conv2d_nhwc(xc2, wc2, bc, stride=(2, 2), padding=(1, 1), act="relu", cache=cache)))
rows.append(("conv3 3x3 s2 [64,100,20,256]", lambda: ops.conv2d_nhwc(xc3, wc2, bc, stride=(2, 2), padding=(1, 1), act="relu", cache=cache)))
for name, fn in rows:
    ts = []
    for cl in ("1", "2", "4"):
        os.environ["APS_B200_TC_CL"] = cl
        ts.append(timeit(fn, 4 if "conv" in name else 20))
    print(f"{name:22s} {ts[0]:8.1f} {ts[1]:8.1f} {ts[2]:8.1f}")
os.environ.pop("APS_B200_TC_CL", None)
