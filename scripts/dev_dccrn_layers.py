"""Per-layer time of the DCCRN convolutions at the bench size (B=128 x 4 s): wraps ops.conv2d_nhwc / conv_transpose2d_nhwc."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aps_b200 import ops  # noqa: E402
from aps_b200.sse.bss import DCCRN  # noqa: E402
from aps_b200.transform import EnhTransform  # noqa: E402

dev = th.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
enh = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, center=True)
net = DCCRN(enh_transform=enh, cplx=True, K="3,3;3,3;3,3;3,3;3,3;3,3;3,3", S="2,1;2,1;2,1;2,1;2,1;2,1;2,1",
            P="1,1,1,1,1,0,0", O="0,0,0,0,0,0,1", C="16,32,64,64,128,128,256", num_spks=2, rnn_resize=512,
            non_linear="sigmoid", connection="cat").to(dev).eval()
x = th.rand(B, 64000, device=dev)
log = []


def wrap(name):
    orig = getattr(ops, name)

    def f(x, w, *a, **k):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        y = orig(x, w, *a, **k)
        e1.record()
        log.append((name, tuple(x.shape), tuple(w.shape), tuple(y.shape), e0, e1))
        return y

    setattr(ops, name, f)


wrap("conv2d_nhwc")
wrap("conv_transpose2d_nhwc")
with th.no_grad():
    net(x)
    log.clear()
    net(x)
th.cuda.synchronize()
tot = 0.0
for name, xs, ws, ys, e0, e1 in log:
    ms = e0.elapsed_time(e1)
    tot += ms
    M = ys[0] * ys[1] * ys[2]
    fl = 2.0 * M * ys[3] * ws[1] * ws[2] * ws[3]
    print(f"{name:>22} x{xs} w{ws} -> y{ys}: {ms:8.3f} ms  {fl/ms/1e9:7.1f} TF/s  M={M} K={ws[1]*ws[2]*ws[3]} N={ys[3]}")
print("total conv ms", tot)
