"""Dev check of the tcgen05 3xTF32 GEMM against an fp64 reference (run on the GPU box)."""
import os
import sys
import time

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["APS_B200_GEMM"] = "tc"
from aps_b200 import ops  # noqa: E402

dev = "cuda:0"
th.manual_seed(0)
for (M, K, N) in ((128, 32, 64), (128, 64, 128), (3200, 256, 2048), (3200, 2048, 256), (3200, 256, 768), (333, 96, 200),
                  (15936, 512, 256), (6400, 2304, 256)):
    x = th.randn(M, K, device=dev)
    w = th.randn(N, K, device=dev) / K**0.5
    b = th.randn(N, device=dev)
    ref = (x.double() @ w.double().t() + b.double())
    ops.GEMM_ENGINE = "tc"
    y = ops.linear(x, w, b)
    th.cuda.synchronize()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    ops.GEMM_ENGINE = "simt"
    y2 = ops.linear(x, w, b)
    err2 = float((y2.double() - ref).abs().max() / ref.abs().max())
    res = {}
    r = th.randn(M, N, device=dev)
    for eng in ("tc", "simt"):
        ops.GEMM_ENGINE = eng
        cache = ops.SplitCache()
        f = lambda: ops.linear(x, w, b, residual=r, cache=cache)
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
        res[eng] = e0.elapsed_time(e1) / 100
    fl = 2.0 * M * K * N
    print(f"M={M} K={K} N={N}: tc err {err:.2e} simt err {err2:.2e} | tc(graph, +split, +res) {res['tc']*1e3:.1f} us "
          f"({fl/res['tc']/1e9:.1f} TF/s eq) simt {res['simt']*1e3:.1f} us ({fl/res['simt']/1e9:.1f} TF/s)", flush=True)
# epilogues on the tc path
ops.GEMM_ENGINE = "tc"
x, w, b, r = th.randn(300, 256, device=dev), th.randn(512, 256, device=dev) / 16, th.randn(512, device=dev), th.randn(300, 512, device=dev)
ref = th.nn.functional.linear(x.double(), w.double(), b.double())
y = ops.linear(x, w, b, act="swish", alpha=0.5, residual=r)
print("swish+res", float((y.double() - (0.5 * ref * th.sigmoid(ref) + r.double())).abs().max()))
wi = th.stack([w[:256], w[256:]], 1).reshape(512, 256).contiguous()
bi = th.stack([b[:256], b[256:]], 1).reshape(512).contiguous()
y = ops.linear(x, wi, bi, act="glu")
print("glu", float((y.double() - th.nn.functional.glu(ref, -1)).abs().max()))
