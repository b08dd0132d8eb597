"""GPU parity of the transform path (kernels F1, F1b, F2, F3, IPD) through the Python shells, i.e.
through the C ABI of libaps_b200.so.  Checked against (a) the committed golden vectors generated from
the live reference, (b) the CPU oracle on seeded inputs, (c) size-independent properties at the
BASELINE.json batch sizes.  Tolerance: integers exact, floats max|d|/max|ref| <= 1e-4 (north_star)."""
import itertools

import pytest
import torch as th

from conftest import FLOAT_TOL, golden_names, load_golden, rel_err
from oracle import transform as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _asr(kw):
    from aps_b200.transform import AsrTransform
    return AsrTransform(**kw).to(DEV).eval()


def _cfg(kw):
    f = O.AsrFeatCfg.__dataclass_fields__
    return O.AsrFeatCfg(**{k: v for k, v in kw.items() if k in f})


# ----------------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("mode", ["librosa", "kaldi", "torch"])
def test_c1_golden(mode):
    """BASELINE config[0]: AsrTransform STFT -> 80-mel fbank, 1 utt 4 s @ 16 kHz."""
    kw, g = load_golden(f"asr_c1_{mode}")
    _, inp = load_golden("asr_c1_input")
    y, n = _asr(kw)(inp["wav"].to(DEV), th.tensor([64000], device=DEV))
    assert th.equal(n.cpu(), g["num_frames"])
    assert y.shape == g["feats"].shape
    assert rel_err(y, g["feats"]) < FLOAT_TOL


@pytest.mark.parametrize("name", golden_names("asr_grid_"))
def test_grid_golden(name):
    kw, g = load_golden(name)
    y, n = _asr(kw)(g["wav"].to(DEV), g["lens"].to(DEV))
    assert th.equal(n.cpu(), g["num_frames"])
    assert rel_err(y, g["feats"]) < FLOAT_TOL


@pytest.mark.parametrize("name", golden_names("stft_")[:-1])
def test_stft_golden(name):
    from aps_b200.transform.utils import STFT
    kw, g = load_golden(name)
    y = STFT(**kw).to(DEV)(g["wav"].to(DEV))
    assert y.shape == g["spec"].shape
    assert rel_err(y, g["spec"]) < FLOAT_TOL


@pytest.mark.parametrize("name", golden_names("istft_"))
def test_istft_golden(name):
    from aps_b200.transform.utils import iSTFT
    kw, g = load_golden(name)
    y = iSTFT(**kw).to(DEV)(g["spec"].to(DEV))
    assert y.shape == g["wav"].shape
    assert rel_err(y, g["wav"]) < FLOAT_TOL


def test_polar_golden():
    from aps_b200.transform.utils import STFT, iSTFT
    kw, g = load_golden("stft_polar")
    pol = STFT(**kw).to(DEV)(g["wav"].to(DEV), return_polar=True)
    assert rel_err(pol[..., 0], g["spec"][..., 0]) < FLOAT_TOL
    # phases: compare on the unit circle, weighted by magnitude (phase of a ~0 bin is arbitrary)
    mag = g["spec"][..., 0].to(DEV)
    d = (th.exp(1j * pol[..., 1]) - th.exp(1j * g["spec"][..., 1].to(DEV))).abs() * mag
    assert float(d.max() / mag.max()) < FLOAT_TOL
    rec = iSTFT(**kw).to(DEV)(g["spec"].to(DEV), return_polar=True)
    assert rel_err(rec, g["rec"]) < FLOAT_TOL


@pytest.mark.parametrize("name", golden_names("enh_"))
def test_enh_golden(name):
    from aps_b200.transform import EnhTransform
    kw, g = load_golden(name)
    t = EnhTransform(**kw).to(DEV).eval()
    lens = th.tensor([4000, 3500], device=DEV)
    packed, n = t.encode(g["wav"].to(DEV), lens)
    assert th.equal(n.cpu(), g["num_frames"])
    assert rel_err(packed, g["packed"]) < FLOAT_TOL
    feats = t(g["packed"].to(DEV))
    assert feats.shape == g["feats"].shape and t.feats_dim == feats.shape[-1]
    assert rel_err(feats, g["feats"]) < FLOAT_TOL
    rec = t.decode([g["packed"][:, 0].to(DEV)])[0]
    assert rel_err(rec, g["rec"]) < FLOAT_TOL


# ----------------------------------------------------------------------------------- oracle, seeded
@pytest.mark.parametrize("mode", ["librosa", "kaldi", "torch"])
def test_features_vs_oracle_sweep(mode):
    th.manual_seed(3)
    x = 0.1 * th.randn(5, 24000)
    x[1] = th.rand(24000)                       # strong DC like the reference's own tests
    x[3, 17000:] = 0                            # zero padded tail (ragged batch)
    lens = th.tensor([24000, 24000, 21000, 17000, 9000])
    for feats, power, lb, pb, an in itertools.product(["fbank-log-cmvn", "spectrogram-log", "fbank"], [False, True],
                                                      [0, 1.0], [True, False], [True, False]):
        kw = dict(feats=feats, stft_mode=mode, use_power=power, log_lower_bound=lb, norm_per_band=pb, audio_norm=an)
        ref, nr = O.AsrFeatures(_cfg(kw))(x.clone(), lens.clone())
        got, ng = _asr(kw)(x.to(DEV), lens.to(DEV))
        assert th.equal(ng.cpu(), nr), kw
        assert rel_err(got, ref) < FLOAT_TOL, kw


@pytest.mark.parametrize("frame_len,frame_hop", [(256, 128), (400, 160), (512, 128), (1024, 256), (100, 50), (400, 133)])
@pytest.mark.parametrize("center", [False, True])
def test_fft_sizes_and_hops(frame_len, frame_hop, center):
    th.manual_seed(4)
    x = 0.1 * th.randn(3, 9000)
    kw = dict(feats="fbank-log-cmvn", frame_len=frame_len, frame_hop=frame_hop, center=center, num_mels=24)
    ref, _ = O.AsrFeatures(_cfg(kw))(x, None)
    got, _ = _asr(kw)(x.to(DEV), None)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < FLOAT_TOL


def test_edge_shapes():
    from aps_b200.transform.utils import STFT
    t = _asr(dict(feats="fbank-log-cmvn"))
    one = th.rand(1, 512, device=DEV)           # exactly one frame
    y, _ = t(one, None)
    assert y.shape == (1, 1, 80)
    ref, _ = O.AsrFeatures(O.AsrFeatCfg())(one.cpu(), None)
    assert rel_err(y, ref) < FLOAT_TOL
    with pytest.raises(RuntimeError):
        t(th.rand(1, 300, device=DEV), None)    # shorter than a frame
    with pytest.raises(AssertionError):
        t(th.rand(2, 4000, device=DEV), th.tensor([4000, 512], device=DEV))   # utils.py:657
    with pytest.raises(RuntimeError, match="2D/3D"):
        STFT(512, 256).to(DEV)(th.rand(2, 2, 2, 4000, device=DEV))
    # strided rows / multi channel views
    big = 0.1 * th.randn(4, 2, 6000, device=DEV)
    ref, _ = O.AsrFeatures(O.AsrFeatCfg())(big.cpu(), None)
    got, _ = t(big, None)
    assert got.shape == ref.shape == (4, 2, 35, 80) and rel_err(got, ref) < FLOAT_TOL
    got2, _ = t(big[:, 1], None)
    assert rel_err(got2, ref[:, 1]) < FLOAT_TOL
    bad = big.clone()
    bad[2, 0, 100] = float("nan")
    with pytest.raises(ValueError, match="NANs"):
        t(bad, None)


def test_emph_side_effect_and_eval_chain():
    """"emph" modifies the caller's waveform in place in the reference (Q9); SpecAug/perturb are identity in eval."""
    kw = dict(feats="perturb-emph-fbank-log-cmvn-aug", pre_emphasis=0.9, aug_prob=1.0)
    th.manual_seed(5)
    x = 0.1 * th.randn(2, 8000)
    xo = x.clone()
    ref, _ = O.AsrFeatures(_cfg(dict(kw, feats="emph-fbank-log-cmvn")))(xo, None)
    xg = x.to(DEV)
    got, _ = _asr(kw)(xg, None)
    assert rel_err(got, ref) < FLOAT_TOL
    expect = x.clone()
    expect[:, 1:] = x[:, 1:] - 0.9 * x[:, :-1]
    assert rel_err(xg, expect) < 1e-6


def test_specaug_training_matches_reference_semantics():
    import random
    from aps_b200.transform.asr import tf_mask
    kw = dict(feats="fbank-log-cmvn-aug", aug_prob=1.0, aug_time_args=(10, 2), aug_freq_args=(8, 2))
    t = _asr(kw).train()
    x = (0.1 * th.randn(3, 8000)).to(DEV)
    base, _ = _asr(dict(kw, feats="fbank-log-cmvn"))(x, None)
    random.seed(11)
    got, _ = t(x, None)
    random.seed(11)
    mask = tf_mask(3, tuple(base.shape[1:]), max_bands=8, max_frame=10, num_freq_masks=2, num_time_masks=2, device=DEV)
    assert th.equal(got, base * mask)


# ----------------------------------------------------------------------------------- properties, full size
def test_c2_full_size_properties():
    """BASELINE config[1]: B=256 x 4 s.  Size-independent checks: frame count, per-frame CMVN moments,
    batch-shard invariance (any row alone == that row inside the batch, bit for bit) and oracle parity on a
    sampled subset of rows."""
    th.manual_seed(0)
    x = 0.1 * th.randn(256, 64000)
    t = _asr(dict(feats="fbank-log-cmvn"))
    lens = th.full((256,), 64000, dtype=th.int64, device=DEV)
    y, n = t(x.to(DEV), lens)
    assert y.shape == (256, 397, 80) and n.tolist() == [397] * 256
    assert float(y.mean(-1).abs().max()) < 1e-4
    assert float(((y**2).mean(-1) - 1).abs().max()) < 1e-3
    rows = [0, 17, 101, 255]
    alone, _ = t(x[rows].to(DEV), None)
    assert th.equal(alone, y[rows])
    ref, _ = O.AsrFeatures(O.AsrFeatCfg())(x[rows], None)
    assert rel_err(y[rows], ref) < FLOAT_TOL


def test_stft_istft_round_trip_full_size():
    """iSTFT(STFT(x)) == x (the reference's test_forward_inverse_stft), at the DCCRN batch (B=128 x 4 s)."""
    from aps_b200.transform.utils import STFT, iSTFT
    th.manual_seed(1)
    x = (0.1 * th.randn(128, 64000)).to(DEV)
    for mode, window in (("librosa", "sqrthann"), ("torch", "hamm"), ("librosa", "hamm")):
        kw = dict(frame_len=512, frame_hop=256, window=window, center=True, mode=mode)
        spec = STFT(**kw).to(DEV)(x)
        assert spec.shape == (128, 257, 251, 2)
        rec = iSTFT(**kw).to(DEV)(spec)
        m = min(rec.shape[-1], x.shape[-1])
        assert rel_err(rec[:, :m], x[:, :m]) < FLOAT_TOL
    # linearity of the transform
    S = STFT(512, 256).to(DEV)
    a, b = x[:8], x[8:16]
    assert rel_err(S(a + 2 * b), S(a) + 2 * S(b)) < 1e-5


def test_enh_multichannel_full_size_subset_vs_oracle():
    """BASELINE config[2] front half: 4-ch STFT of B=64 x 4 s; parity on sampled rows."""
    from aps_b200.transform import EnhTransform
    th.manual_seed(2)
    x = 0.1 * th.randn(64, 4, 64000)
    t = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann").to(DEV)
    packed, n = t.encode(x.to(DEV), th.full((64,), 64000, device=DEV))
    assert packed.shape == (64, 4, 257, 249, 2) and n.tolist() == [249] * 64
    K, w = O.dft_kernel(512, O.window("sqrthann", 512))
    rows = [0, 31, 63]
    ref = O.stft_dense(x[rows], K, w, 256)
    assert rel_err(packed[rows], ref) < FLOAT_TOL
    feats = t(packed)
    reff = O.cmvn(O.log_compress(O.magnitude(ref[:, 0]).transpose(-1, -2)))
    assert feats.shape == (64, 249, 257)
    assert rel_err(feats[rows], reff) < FLOAT_TOL


def test_missing_library_fails_loudly(monkeypatch):
    from aps_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libaps_b200.so")
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()
