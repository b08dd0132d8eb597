"""Drop-in proof: the reference's own factories (`aps.libs.aps_transform`,
`aps_asr_nnet`, `aps_sse_nnet`, `aps_task`) hand out this package's modules after `aps_b200.register_into_aps()`, the
networks built that way keep the reference's `state_dict` layout (strict loading in both directions), and — on the GPU —
the reference's `CtcASR._training_prep` (aps/asr/ctc.py:113-134) runs on this transform and encoder and reproduces a
golden computed by the unmodified reference.

The registry is process-global state, so the take-over runs in a child interpreter.  The CPU test needs the live
reference tree (marker `reference`); the GPU test uses the copy staged in oracle/_ref by oracle/build_ref.sh (it travels to
the GPU box with the snapshot) and is skipped when that is absent."""
import json
import os
import subprocess
import sys

import pytest

from conftest import FLOAT_TOL, GOLDEN, REFERENCE, ROOT

CHILD_CPU = r'''
import copy, json, sys, warnings
warnings.filterwarnings("ignore")
sys.path[:0] = [sys.argv[1], sys.argv[2], sys.argv[3]]          # repo root, shims, reference
import torch as th
from aps.libs import aps_transform, aps_asr_nnet, aps_sse_nnet, aps_task, ApsRegisters

ASR_T = dict(feats="fbank-log-cmvn", frame_len=400, frame_hop=160, window="hamm", pre_emphasis=0.97, num_mels=80)
ENH_T = dict(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, center=True)
ENC = dict(arch_kwargs=dict(att_dim=128, nhead=4, feedforward_dim=256, att_dropout=0.1, ffn_dropout=0.1, kernel_size=15),
           num_layers=2, proj="conv2d", proj_kwargs=dict(conv_channels=32, num_layers=2), pose="rel",
           pose_kwargs=dict(lradius=16, rradius=16))
NETS = {
    "sse@dccrn": dict(cplx=True, K="3,3;3,3;3,3;3,3;3,3;3,3;3,3", S="2,1;2,1;2,1;2,1;2,1;2,1;2,1", P="1,1,1,1,1,0,0",
                      O="0,0,0,0,0,0,1", C="16,32,64,64,128,128,256", num_spks=2, rnn_resize=512, non_linear="sigmoid",
                      connection="cat"),
    "sse@freq_tcn": dict(in_features=257, num_bins=257, num_spks=1, non_linear="sigmoid", N=2, B=2),
    "sse@time_tcn": dict(L=20, N=64, X=2, R=2, B=64, H=128, P=3, num_spks=2),
    "sse@freq_xfmr": dict(input_size=257, num_spks=2, num_bins=257, arch="xfmr", pose="rel", num_layers=2,
                          arch_kwargs=dict(att_dim=128, nhead=4, feedforward_dim=256), pose_kwargs=dict(lradius=8, rradius=8)),
}

def build(kind):
    out = {}
    tf = aps_transform("asr")(**ASR_T)
    out["asr@ctc"] = aps_asr_nnet("asr@ctc")(input_size=80, vocab_size=40, ctc=True, ead=False, asr_transform=tf, enc_type="cfmr",
                                             enc_kwargs=copy.deepcopy(ENC))
    for name, kw in NETS.items():
        enh = aps_transform("enh")(**ENH_T)
        out[name] = aps_sse_nnet(name)(**copy.deepcopy(kw)) if name == "sse@time_tcn" else aps_sse_nnet(name)(enh_transform=enh, **copy.deepcopy(kw))
    return out

ref = build("ref")
ref_cls = {k: type(v).__module__ for k, v in ref.items()}
import aps_b200
done = aps_b200.register_into_aps()
ours = build("ours")
report = {"registered": sorted(done), "nets": {}}
for name in ref:
    a, b = ref[name], ours[name]
    sa, sb = a.state_dict(), b.state_dict()
    same_keys = list(sa.keys()) == list(sb.keys())
    same_shapes = same_keys and all(sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype for k in sa)
    b.load_state_dict(sa, strict=True)          # reference checkpoint -> ours
    a.load_state_dict(sb, strict=True)          # ours -> reference
    mods = sorted({type(m).__module__.split(".")[0] for m in b.modules()} - {"torch"})
    report["nets"][name] = dict(same_keys=same_keys, same_shapes=same_shapes, ref_module=ref_cls[name],
                                module=type(b).__module__, packages=mods, nparams=len(sa))
net = ours["asr@ctc"]
report["ctc_encoder"] = type(net.encoder).__module__
report["ctc_transform"] = type(net.asr_transform).__module__
task = aps_task("sse@sisnr", ours["sse@dccrn"], num_spks=2)
report["task"] = type(task).__module__
report["transform_asr"] = ApsRegisters.transform["asr"].__module__
report["transform_enh"] = ApsRegisters.transform["enh"].__module__
print("REPORT " + json.dumps(report))
'''


@pytest.mark.reference
def test_reference_factories_hand_out_this_package_and_checkpoints_load_both_ways(tmp_path):
    script = tmp_path / "child.py"
    script.write_text(CHILD_CPU)
    out = subprocess.run([sys.executable, str(script), ROOT, os.path.join(ROOT, "oracle", "ref_shims"), REFERENCE],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    rep = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("REPORT ")][0][7:])
    assert rep["transform_asr"].startswith("aps_b200") and rep["transform_enh"].startswith("aps_b200")
    assert rep["ctc_encoder"].startswith("aps_b200") and rep["ctc_transform"].startswith("aps_b200")
    assert rep["task"].startswith("aps_b200")
    for name, r in rep["nets"].items():
        assert r["same_keys"] and r["same_shapes"], (name, r)
        assert r["ref_module"].startswith("aps.")
        if name == "asr@ctc":
            assert r["module"].startswith("aps.")           # the reference's CtcASR shell ...
            assert "aps_b200" in r["packages"]               # ... over this package's transform and encoder
        else:
            assert r["module"].startswith("aps_b200"), (name, r)
    for alias in ("sse:sse@dccrn", "sse:sse@freq_tcn", "sse:sse@time_tcn", "sse:sse@freq_xfmr", "task:sse@sisnr",
                  "transform:asr", "transform:enh", "aps.asr.ctc.TransformerEncoder"):
        assert alias in rep["registered"], alias


CHILD_GPU = r'''
import copy, json, sys, warnings
warnings.filterwarnings("ignore")
sys.path[:0] = [sys.argv[1], sys.argv[2]]                        # repo root, staged reference (aps + shims)
import numpy as np, torch as th
import aps_b200
from aps.libs import aps_transform, aps_asr_nnet
aps_b200.register_into_aps()
z = np.load(sys.argv[3], allow_pickle=False)
kw = json.loads(str(z["kwargs"]))
tf = aps_transform("asr")(**kw["transform"])
net = aps_asr_nnet("asr@ctc")(asr_transform=tf, **copy.deepcopy(kw["net"]))
assert type(net).__module__ == "aps.asr.ctc" and type(net.encoder).__module__.startswith("aps_b200")
net.encoder.load_state_dict({k[2:]: th.from_numpy(z[k]) for k in z.files if k.startswith("p.")}, strict=True)
net = net.to("cuda:0").eval()
x, lens = th.from_numpy(z["x"]).to("cuda:0"), th.from_numpy(z["lens"])
with th.no_grad():
    enc_out, enc_ctc, enc_len = net(x, lens)                     # CtcASR.forward -> _training_prep (aps/asr/ctc.py:113-134)
ref = th.from_numpy(z["enc_out"])
err = float((enc_out.cpu() - ref).abs().max() / ref.abs().max())
print("REPORT " + json.dumps(dict(err=err, shape=list(enc_out.shape), len=enc_len.tolist(), ref_len=z["enc_len"].tolist(),
                                  fast=bool(net.encoder._fast))))
'''


@pytest.mark.gpu
def test_reference_ctc_training_prep_on_this_transform_and_encoder_gpu(tmp_path):
    staged = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(staged, "aps")):
        pytest.skip("oracle/_ref not staged (run oracle/build_ref.sh in the build container)")
    script = tmp_path / "child.py"
    script.write_text(CHILD_GPU)
    out = subprocess.run([sys.executable, str(script), ROOT, staged, os.path.join(GOLDEN, "ctc_0.npz")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    rep = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("REPORT ")][0][7:])
    assert rep["len"] == rep["ref_len"]                          # integer exact
    assert rep["err"] < FLOAT_TOL, rep
    assert rep["fast"], "att_dim 128: the TMA-fed pair path must be the one that ran"
