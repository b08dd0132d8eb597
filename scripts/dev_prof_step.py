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
elif wl == "mvdr_tcn":
    from aps_b200.asr.filter import MvdrBeamformer
    from aps_b200.cplx import ComplexTensor
    from aps_b200.sse.bss import FreqConvTasNet
    from aps_b200.transform import EnhTransform
    B, C = 64, 4
    enh = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann")
    tcn = FreqConvTasNet(enh_transform=enh, in_features=257, num_bins=257, num_spks=1, non_linear="sigmoid").to(dev).eval()
    mvdr = MvdrBeamformer(257, att_dim=512).to(dev).eval()
    x = 0.1 * th.randn(B, C, S, device=dev)

    def step():
        packed, _ = tcn.enh_transform.encode(x, None)
        mask = tcn.mask_predict(tcn.enh_transform(packed))
        return mvdr(mask.transpose(1, 2), ComplexTensor(packed[..., 0], packed[..., 1]))
elif wl == "encoder":
    import copy
    from aps_b200.asr.transformer import TransformerEncoder
    cfg = dict(arch="cfmr", input_size=80, output_proj=-1, num_layers=12, proj="conv2d",
               proj_kwargs=dict(conv_channels=256, num_layers=3), pose="rel",
               pose_kwargs=dict(dropout=0.1, lradius=256, rradius=256),
               arch_kwargs=dict(att_dim=256, nhead=4, feedforward_dim=2048, att_dropout=0.1, ffn_dropout=0.1,
                                kernel_size=15, pre_norm=False))
    net = TransformerEncoder(**copy.deepcopy(cfg)).to(dev).eval()
    net.use_graphs = False
    x = th.randn(64, 398, 80, device=dev)
    step = lambda: net(x, None)
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
