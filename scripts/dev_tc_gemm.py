"""Dev check of the persistent tcgen05 3xTF32 GEMM / implicit convolutions (run on the GPU box):
errors against fp64 / torch references for every tile width, then CUDA-graph timings vs the SIMT engine."""
import os
import sys

import torch as th
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["APS_B200_GEMM"] = "tc"
from aps_b200 import ops  # noqa: E402

dev = "cuda:0"
th.manual_seed(0)
quick = "--quick" in sys.argv


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


# ---- correctness: linear, all tile widths, ragged shapes, more tiles than SMs -------------------------------
for bn in ("64", "128", "256", ""):
    os.environ["APS_B200_TC_BN"] = bn
    for (M, K, N) in ((128, 32, 64), (333, 96, 200), (300, 64, 257), (3200, 256, 768), (700, 2304, 256), (40000, 64, 320)):
        x = th.randn(M, K, device=dev)
        w = th.randn(N, K, device=dev) / K**0.5
        b = th.randn(N, device=dev)
        r = th.randn(M, N, device=dev)
        ref = x.double() @ w.double().t() + b.double()
        y = ops.linear(x, w, b)
        y2 = ops.linear(x, w, b, act="swish", alpha=0.5, residual=r)
        ref2 = 0.5 * ref * th.sigmoid(ref) + r.double()
        th.cuda.synchronize()
        print(f"BN={bn or 'auto':>4} linear M={M} K={K} N={N}: err {relerr(y, ref):.2e} swish+res {relerr(y2, ref2):.2e}", flush=True)
    x, w, b = th.randn(300, 256, device=dev), th.randn(512, 256, device=dev) / 16, th.randn(512, device=dev)
    ref = F.linear(x.double(), w.double(), b.double())
    wi = th.stack([w[:256], w[256:]], 1).reshape(512, 256).contiguous()
    bi = th.stack([b[:256], b[256:]], 1).reshape(512).contiguous()
    print(f"BN={bn or 'auto':>4} glu err {relerr(ops.linear(x, wi, bi, act='glu'), F.glu(ref, -1)):.2e}", flush=True)
    ref = F.linear(x.double(), w[:258].double(), b[:258].double())
    wi = th.stack([w[:129], w[129:258]], 1).reshape(258, 256).contiguous()
    bi = th.stack([b[:129], b[129:258]], 1).reshape(258).contiguous()
    print(f"BN={bn or 'auto':>4} glu(unaligned) err {relerr(ops.linear(x, wi, bi, act='glu'), F.glu(ref, -1)):.2e}", flush=True)
    sl, ps, pt = th.rand(512, device=dev), th.rand(512, device=dev) + 0.5, th.randn(512, device=dev)
    ref = F.linear(x.double(), w.double(), b.double())
    refp = th.where(ref >= 0, ref, ref * sl.double()) * ps.double() + pt.double()
    print(f"BN={bn or 'auto':>4} prelu+affine err {relerr(ops.linear(x, w, b, act='prelu', slope=sl, post=(ps, pt)), refp):.2e}", flush=True)
    # implicit conv / transposed conv vs torch (fp64 on the CPU)
    for (B, H, W, Ci, Co, k, s_, p_, d_) in ((3, 40, 21, 32, 64, (3, 3), (2, 2), (1, 1), (1, 1)),
                                             (1, 18, 9, 64, 32, (5, 2), (2, 1), (2, 0), (1, 1)),
                                             (2, 31, 17, 32, 288, (3, 3), (2, 1), (0, 1), (1, 2)),
                                             (4, 100, 20, 256, 256, (3, 3), (2, 2), (1, 1), (1, 1))):
        x, w, b = th.randn(B, Ci, H, W), th.randn(Co, Ci, *k) * 0.1, th.randn(Co)
        ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), stride=s_, padding=p_, dilation=d_), 0.01).permute(0, 2, 3, 1)
        got = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(dev), w.permute(0, 2, 3, 1).contiguous().to(dev),
                              b.to(dev), stride=s_, padding=p_, dilation=d_, act="leaky_relu", leaky=0.01)
        print(f"BN={bn or 'auto':>4} conv {B}x{H}x{W}x{Ci}->{Co} k{k} s{s_}: err {relerr(got.cpu(), ref):.2e}", flush=True)
    for (B, H, W, Ci, Co, k, s_, p_, op_) in ((2, 9, 30, 64, 32, (3, 3), (2, 1), (1, 1), (0, 0)),
                                              (2, 4, 25, 256, 128, (3, 3), (2, 1), (0, 1), (1, 0)),
                                              (1, 17, 11, 32, 64, (3, 3), (2, 2), (1, 1), (1, 1)),
                                              (3, 17, 30, 64, 4, (3, 3), (2, 1), (1, 1), (0, 0)),
                                              (2, 7, 13, 32, 36, (3, 3), (1, 1), (1, 1), (0, 0)),
                                              (2, 6, 9, 32, 40, (5, 3), (3, 1), (2, 1), (2, 0)),
                                              (5, 40, 60, 32, 32, (3, 3), (2, 1), (0, 1), (1, 0))):
        x, w, b = th.randn(B, Ci, H, W), th.randn(Ci, Co, *k) * 0.1, th.randn(Co)
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=s_, padding=p_, output_padding=op_).permute(0, 2, 3, 1)
        got = ops.conv_transpose2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(dev),
                                        w.transpose(0, 1).permute(0, 2, 3, 1).contiguous().to(dev), b.to(dev), stride=s_,
                                        padding=p_, output_padding=op_)
        print(f"BN={bn or 'auto':>4} tconv {B}x{H}x{W}x{Ci}->{Co} s{s_}: err {relerr(got.cpu(), ref):.2e}", flush=True)
os.environ["APS_B200_TC_BN"] = ""
if quick:
    sys.exit(0)


# ---- timings (CUDA graph of 20 calls, residual epilogue) ------------------------------------------------------
def timed(f):
    for _ in range(3):
        f()
    th.cuda.synchronize()
    g = th.cuda.CUDAGraph()
    with th.cuda.graph(g):
        for _ in range(20):
            f()
    g.replay()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / 100


for (M, K, N) in ((3200, 256, 2048), (3200, 2048, 256), (3200, 256, 768), (3200, 256, 256), (15936, 512, 256),
                  (15936, 256, 512), (128000, 2304, 256), (32000, 2304, 256)):
    x = th.randn(M, K, device=dev)
    w = th.randn(N, K, device=dev) / K**0.5
    b = th.randn(N, device=dev)
    r = th.randn(M, N, device=dev)
    fl = 2.0 * M * K * N
    line = f"M={M} K={K} N={N}:"
    for eng, bn in (("tc", "64"), ("tc", "128"), ("tc", "256"), ("tc", ""), ("simt", "")):
        if eng == "simt" and M * K * N > 4e10:
            continue
        ops.GEMM_ENGINE = eng
        os.environ["APS_B200_TC_BN"] = bn
        cache = ops.SplitCache()
        t = timed(lambda: ops.linear(x, w, b, residual=r, cache=cache))
        line += f"  {eng}{bn or ''} {t*1e3:.1f} us ({fl/t/1e9:.1f} TF/s)"
    print(line, flush=True)
ops.GEMM_ENGINE = "tc"
os.environ["APS_B200_TC_BN"] = ""
for (B, H, W, Ci, Co, s_) in ((64, 199, 40, 256, 256, (2, 2)), (64, 100, 20, 256, 256, (2, 2)), (128, 129, 251, 32, 64, (2, 1)),
                              (128, 33, 251, 128, 128, (2, 1))):
    x = th.randn(B, H, W, Ci, device=dev)
    w = th.randn(Co, 3, 3, Ci, device=dev) * 0.05
    b = th.randn(Co, device=dev)
    line = f"conv {B}x{H}x{W}x{Ci}->{Co} s{s_}:"
    for eng in ("tc", "simt"):
        ops.GEMM_ENGINE = eng
        cache = ops.SplitCache()
        out = ops.conv2d_nhwc(x, w, b, stride=s_, padding=(1, 1), act="relu", cache=cache)
        fl = 2.0 * out.numel() * 9 * Ci
        t = timed(lambda: ops.conv2d_nhwc(x, w, b, stride=s_, padding=(1, 1), act="relu", cache=cache))
        line += f"  {eng} {t*1e3:.1f} us ({fl/t/1e9:.1f} TF/s)"
    print(line, flush=True)
