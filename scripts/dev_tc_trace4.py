"""Timeline (trace build) of CTA 0 for a small-N transposed convolution of the DCCRN decoder."""
import ctypes
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import _lib, ops  # noqa: E402

NAMES = {1: "tma_issue", 2: "mma_start", 3: "mma_tile_commit", 4: "epi_start", 5: "epi_tmem_release", 6: "epi_done",
         7: "prod_stored", 8: "prod_slot_free", 9: "kernel_start", 11: "mma_a_ready"}
dev = "cuda:0"
lib = _lib.load()
lib.aps_b200_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
B, H, W, Ci, Co = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (16, 129, 251, 64, 4)))
x = th.randn(B, H, W, Ci, device=dev)
w = th.randn(Co, 3, 3, Ci, device=dev) * 0.05
b = th.randn(Co, device=dev)
cache = ops.SplitCache()
f = lambda: ops.conv_transpose2d_nhwc(x, w, b, stride=(2, 1), padding=(1, 1), output_padding=(0, 0), cache=cache)
for _ in range(3):
    f()
buf = th.zeros(1024, dtype=th.int64, device=dev)
lib.aps_b200_tc_trace(buf.data_ptr(), 1024)
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record()
f()
e1.record()
th.cuda.synchronize()
lib.aps_b200_tc_trace(None, 0)
h = buf.cpu().tolist()
n = min(h[0] & 0xFFFFFFFF, 1022)
ev = sorted(((v & ((1 << 48) - 1)), (v >> 48) & 0xFFFF) for v in h[1:1 + n])
t0 = ev[0][0]
print(f"wall {e0.elapsed_time(e1)*1e3:.0f} us, {n} events; producer cycles: wait {h[1020]} store {h[1021]} gather {h[1022]} (seek {h[1017]} set_tile {h[1018]} loads {h[1019]})")
for t, e in ev[:40]:
    print(f"{t - t0:>8} {NAMES.get(e, e)}")
