"""MVDR front-end (rows a14-a18): oracle vs golden/live reference on the CPU, CUDA kernels vs both on the GPU."""
import pytest
import torch as th

from conftest import FLOAT_TOL, golden_names, import_reference, load_golden, rel_err
from oracle import mvdr as OM

DEV = "cuda:0"


def _params(g):
    return {k[2:]: v for k, v in g.items() if k.startswith("p.")}


@pytest.mark.parametrize("name", golden_names("mvdr_"))
def test_oracle_mvdr_golden(name):
    kw, g = load_golden(name)
    y = OM.mvdr_forward(g["mask_s"], (g["xr"], g["xi"]), _params(g), mask_n=g["mask_n"] if kw["use_n"] else None,
                        x_len=g["lens"] if kw["use_len"] else None, mask_norm=kw["mask_norm"])
    assert rel_err(y[0], g["yr"]) < 1e-5 and rel_err(y[1], g["yi"]) < 1e-5
    R = OM.estimate_covar(g["mask_s"].transpose(1, 2), (g["xr"], g["xi"]))
    assert rel_err(R[0], g["Rr"]) < 1e-6 and rel_err(R[1], g["Ri"]) < 1e-6


@pytest.mark.reference
def test_oracle_mvdr_vs_live_reference():
    import_reference()
    from aps.asr.filter.mvdr import MvdrBeamformer
    from aps.cplx import ComplexTensor
    th.manual_seed(0)
    N, C, F, T = 3, 4, 65, 40
    net = MvdrBeamformer(F, att_dim=32).eval()
    xr, xi, ms, mn = th.randn(N, C, F, T), th.randn(N, C, F, T), th.rand(N, T, F), th.rand(N, T, F)
    lens = th.tensor([40, 33, 20])
    for use_n, use_len in ((False, False), (True, True), (False, True)):
        with th.no_grad():
            y = net(ms, ComplexTensor(xr, xi), mask_n=mn if use_n else None, x_len=lens if use_len else None)
        o = OM.mvdr_forward(ms, (xr, xi), dict(net.state_dict()), mask_n=mn if use_n else None,
                            x_len=lens if use_len else None)
        assert rel_err(o[0], y.real) < 1e-6 and rel_err(o[1], y.imag) < 1e-6


def test_state_dict_keys():
    from aps_b200.asr.filter import MvdrBeamformer
    assert list(MvdrBeamformer(257, att_dim=512).state_dict()) == ["ref.proj.weight", "ref.proj.bias",
                                                                   "ref.gvec.weight", "ref.gvec.bias"]


def _net(kw, g):
    from aps_b200.asr.filter import MvdrBeamformer
    net = MvdrBeamformer(kw["num_bins"], att_dim=kw["att_dim"], mask_norm=kw["mask_norm"])
    net.load_state_dict(_params(g), strict=True)
    return net.to(DEV).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names("mvdr_"))
def test_mvdr_golden_gpu(name):
    from aps_b200.asr.filter import estimate_covar
    from aps_b200.cplx import ComplexTensor
    kw, g = load_golden(name)
    x = ComplexTensor(g["xr"].to(DEV), g["xi"].to(DEV))
    y = _net(kw, g)(g["mask_s"].to(DEV), x, mask_n=g["mask_n"].to(DEV) if kw["use_n"] else None,
                    x_len=g["lens"].to(DEV) if kw["use_len"] else None)
    assert isinstance(y, ComplexTensor) and y.real.shape == g["yr"].shape
    assert rel_err(y.real, g["yr"]) < FLOAT_TOL and rel_err(y.imag, g["yi"]) < FLOAT_TOL
    R = estimate_covar(g["mask_s"].to(DEV).transpose(1, 2), x)
    assert rel_err(R.real, g["Rr"]) < FLOAT_TOL and rel_err(R.imag, g["Ri"]) < FLOAT_TOL


@pytest.mark.gpu
def test_mvdr_full_size_vs_oracle_subset():
    """BASELINE config[2] back half: 4-ch MVDR on B=64 x 4 s (T=249, F=257) fed by the packed STFT views;
    parity on sampled utterances + the distortionless property w^H steering = 1 is replaced by the
    size-independent identity sum_c u_c = 1 and batch-shard invariance."""
    from aps_b200.asr.filter import MvdrBeamformer, beamform
    from aps_b200.cplx import ComplexTensor
    from aps_b200.transform import EnhTransform
    th.manual_seed(5)
    N, C = 64, 4
    wav = 0.1 * th.randn(N, C, 64000)
    enh = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256).to(DEV)
    packed, nf = enh.encode(wav.to(DEV), th.full((N,), 64000, device=DEV))
    x = ComplexTensor(packed[..., 0], packed[..., 1])          # strided halves of the packed STFT
    T, F = packed.shape[3], packed.shape[2]
    mask = th.rand(N, T, F)
    lens = th.full((N,), T, dtype=th.int64)
    lens[1::2] = T - 40
    net = MvdrBeamformer(F, att_dim=512).to(DEV).eval()
    y = net(mask.to(DEV), x, x_len=lens.to(DEV))
    assert y.real.shape == (N, T, F)
    rows = [0, 1, 63]
    pc = packed[rows].cpu()
    ref = OM.mvdr_forward(mask[rows], (pc[..., 0], pc[..., 1]), {k: v.cpu() for k, v in net.state_dict().items()},
                          x_len=lens[rows])
    assert rel_err(y.real[rows], ref[0]) < FLOAT_TOL and rel_err(y.imag[rows], ref[1]) < FLOAT_TOL
    alone = net(mask[rows].to(DEV), ComplexTensor(packed[rows][..., 0], packed[rows][..., 1]), x_len=lens[rows].to(DEV))
    assert th.equal(alone.real, y.real[rows]) and th.equal(alone.imag, y.imag[rows])
    # beamform() with unit weight on channel 2 returns channel 2
    w = th.zeros(N, C, F, device=DEV)
    w[:, 2] = 1
    b = beamform(ComplexTensor(w, th.zeros_like(w)), x)
    assert th.equal(b.real, packed[:, 2, ..., 0]) and th.equal(b.imag, packed[:, 2, ..., 1])


@pytest.mark.gpu
def test_mvdr_errors():
    from aps_b200.asr.filter import MvdrBeamformer
    from aps_b200.cplx import ComplexTensor
    net = MvdrBeamformer(17, att_dim=8).to(DEV)
    with pytest.raises(RuntimeError, match="2..6 channels"):
        net(th.rand(1, 5, 17, device=DEV), ComplexTensor(th.rand(1, 7, 17, 5, device=DEV)))
    with pytest.raises(RuntimeError, match="complex"):
        net(th.rand(1, 5, 17, device=DEV), th.rand(1, 4, 17, 5, device=DEV))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(th.rand(1, 5, 17), ComplexTensor(th.rand(1, 4, 17, 5)))
