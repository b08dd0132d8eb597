"""Frequency-domain Conv-TasNet mask estimator (row a19) and the MVDR + TCN composition of BASELINE config[2]."""
import pytest
import torch as th

from conftest import FLOAT_TOL, golden_names, load_golden, rel_err
from oracle import mvdr as OM
from oracle import tcn as OT

DEV = "cuda:0"


def _sd(g, strip_enh=False):
    sd = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    return {k: v for k, v in sd.items() if not k.startswith("enh_transform")} if strip_enh else sd


def _oracle_masks(cfg, g):
    n = cfg["net"]
    return OT.tf_mask(_sd(g, True), g["feats"], n["N"], n["B"], n["num_spks"], n.get("norm", "BN"), n["non_linear"],
                      n.get("causal", False), n.get("skip_residual", False))


@pytest.mark.parametrize("name", golden_names("tcn_"))
def test_oracle_tcn_golden(name):
    cfg, g = load_golden(name)
    m = th.stack(_oracle_masks(cfg, g))
    ref = g["masks"] if g["masks"].dim() == 4 else g["masks"][None]
    assert rel_err(m, ref) < 1e-5


def _net(cfg, g):
    from aps_b200.sse.bss import FreqConvTasNet
    from aps_b200.transform import EnhTransform
    net = FreqConvTasNet(enh_transform=EnhTransform(**cfg["enh"]), **cfg["net"])
    sd = _sd(g)
    sd.update({k: v for k, v in net.state_dict().items() if k.endswith(".K") and k not in sd})   # DFT matrices are not stored
    net.load_state_dict(sd, strict=True)
    return net


@pytest.mark.parametrize("name", golden_names("tcn_"))
def test_tcn_state_dict_loads_strict(name):
    cfg, g = load_golden(name)
    _net(cfg, g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names("tcn_"))
def test_tcn_golden_gpu(name):
    cfg, g = load_golden(name)
    net = _net(cfg, g).to(DEV).eval()
    masks = net.mask_predict(g["feats"].to(DEV))
    assert masks.shape == g["masks"].shape
    assert rel_err(masks, g["masks"]) < FLOAT_TOL
    net.training_mode = "time"                                  # STFT -> feats -> masks -> masked iSTFT
    wav = net(g["mix"].to(DEV))
    wav = th.stack(wav) if isinstance(wav, list) else wav
    assert wav.shape == g["wav"].shape and rel_err(wav, g["wav"]) < FLOAT_TOL
    one = net.infer(g["mix"][0].to(DEV), mode="time")
    one = th.stack(one) if isinstance(one, list) else one
    ref1 = g["wav"][:, 0] if g["wav"].dim() == 3 else g["wav"][0]
    assert rel_err(one, ref1) < FLOAT_TOL


@pytest.mark.gpu
def test_c3_mvdr_plus_tcn_full_size():
    """BASELINE config[2]: 4-ch STFT -> ref-channel log-spectrogram-cmvn -> freq-TCN sigmoid mask ->
    MVDR (covariance + solve + beamform), B = 64 x 4 s.  Parity of sampled utterances against the oracles."""
    from aps_b200.asr.filter import MvdrBeamformer
    from aps_b200.cplx import ComplexTensor
    from aps_b200.sse.bss import FreqConvTasNet
    from aps_b200.transform import EnhTransform
    from oracle import transform as O
    th.manual_seed(7)
    N, C = 64, 4
    wav = 0.1 * th.randn(N, C, 64000)
    enh = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann")
    tcn = FreqConvTasNet(enh_transform=enh, in_features=257, num_bins=257, num_spks=1, non_linear="sigmoid").eval()
    mvdr = MvdrBeamformer(257, att_dim=512).eval()
    with th.no_grad():                       # non-trivial BatchNorm statistics
        for name, buf in tcn.named_buffers():
            if name.endswith("running_var"):
                buf.copy_(0.5 + th.rand(buf.shape))
    tcn_d, mvdr_d = tcn.to(DEV), mvdr.to(DEV)
    packed, _ = tcn_d.enh_transform.encode(wav.to(DEV), None)
    feats = tcn_d.enh_transform(packed)
    mask = tcn_d.mask_predict(feats)                                        # N x F x T (views)
    y = mvdr_d(mask.transpose(1, 2), ComplexTensor(packed[..., 0], packed[..., 1]))
    assert mask.shape == (N, 257, 249) and y.real.shape == (N, 249, 257)
    rows = [0, 63]
    K, w = O.dft_kernel(512, O.window("sqrthann", 512))
    pr = O.stft_dense(wav[rows], K, w, 256)
    fr = O.cmvn(O.log_compress(O.magnitude(pr[:, 0]).transpose(-1, -2)))
    sd = {k: v.cpu() for k, v in tcn_d.state_dict().items() if not k.startswith("enh_transform")}
    mr = OT.tf_mask(sd, fr, 3, 6, 1, "BN", "sigmoid")[0]
    assert rel_err(mask[rows], mr) < FLOAT_TOL
    yr = OM.mvdr_forward(mr.transpose(1, 2), (pr[..., 0], pr[..., 1]), {k: v.cpu() for k, v in mvdr_d.state_dict().items()})
    assert rel_err(y.real[rows], yr[0]) < 5 * FLOAT_TOL and rel_err(y.imag[rows], yr[1]) < 5 * FLOAT_TOL
