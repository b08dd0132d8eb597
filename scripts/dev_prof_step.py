"""Per-kernel device time of one bench workload step with torch.profiler (cheap alternative to an ncu launch list)."""
import os
import sys

import torch as th
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
wl = sys.argv[1] if len(sys.argv) > 1 else "dccrn"
dev = th.device("cuda", 0)
S = 64000
if wl == "dccrn":
    from aps_b200.sse.bss import DCCRN
    from aps_b200.task import SisnrTask
    from aps_b200.transform import EnhTransform
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    enh = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, center=True)
    net = DCCRN(enh_transform=enh, cplx=True, K="3,3;3,3;3,3;3,3;3,3;3,3;3,3", S="2,1;2,1;2,1;2,1;2,1;2,1;2,1",
                P="1,1,1,1,1,0,0", O="0,0,0,0,0,0,1", C="16,32,64,64,128,128,256", num_spks=2, rnn_resize=512,
                non_linear="sigmoid", connection="cat").to(dev).eval()
    task = SisnrTask(net, num_spks=2)
    x = th.rand(B, S, device=dev)
    egs = {"mix": x, "ref": [0.5 * x, 0.5 * x.flip(-1)]}
    step = lambda: task(egs)["loss"]
else:
    raise SystemExit("unknown workload")
with th.no_grad():
    for _ in range(2):
        step()
    th.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        th.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
