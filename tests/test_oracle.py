"""The oracle (oracle/transform.py) against (a) the committed golden vectors generated from the live
reference and (b) the live reference itself when its tree is present (build container)."""
import itertools

import pytest
import torch as th

from conftest import golden_names, import_reference, load_golden, rel_err
from oracle import transform as O


def _asr_cfg(kw):
    fields = O.AsrFeatCfg.__dataclass_fields__
    return O.AsrFeatCfg(**{k: v for k, v in kw.items() if k in fields})


@pytest.mark.parametrize("mode", ["librosa", "kaldi", "torch"])
def test_oracle_c1_golden(mode):
    kw, g = load_golden(f"asr_c1_{mode}")
    _, inp = load_golden("asr_c1_input")
    y, n = O.AsrFeatures(_asr_cfg(kw))(inp["wav"], th.tensor([64000]))
    assert th.equal(n, g["num_frames"])                     # integers: exact
    assert rel_err(y, g["feats"]) < 1e-5


@pytest.mark.parametrize("name", [n for n in golden_names("asr_grid_") if n[-1] in "012345"])
def test_oracle_grid_golden(name):
    kw, g = load_golden(name)
    y, n = O.AsrFeatures(_asr_cfg(kw))(g["wav"], g["lens"])
    assert th.equal(n, g["num_frames"])
    assert rel_err(y, g["feats"]) < 1e-5


@pytest.mark.parametrize("name", golden_names("stft_")[:-1])
def test_oracle_stft_golden(name):
    kw, g = load_golden(name)
    mode = kw["mode"]
    w0 = O.window(kw["window"], kw["frame_len"])
    if mode == "torch":
        nfft = O.fft_size_of(kw["frame_len"])
        y = O.stft_torch(g["wav"], kw["frame_len"], kw["frame_hop"], w0, nfft, center=kw["center"])
        y = y.reshape(g["spec"].shape)                      # reference quirk: [N*C, 1, ...] for 3-D input
    else:
        K, w = O.dft_kernel(kw["frame_len"], w0, mode=mode)
        y = O.stft_dense(g["wav"], K, w, kw["frame_hop"], center=kw["center"])
    assert rel_err(y, g["spec"]) < 1e-5


@pytest.mark.parametrize("name", golden_names("istft_"))
def test_oracle_istft_golden(name):
    kw, g = load_golden(name)
    w0 = O.window(kw["window"], kw["frame_len"])
    if kw["mode"] == "torch":
        y = O.istft_torch(g["spec"], kw["frame_hop"], w0, O.fft_size_of(kw["frame_len"]), center=kw["center"])
    else:
        K, w = O.dft_kernel(kw["frame_len"], w0, inverse=True, mode=kw["mode"])
        y = O.istft_dense(g["spec"], K, w, kw["frame_hop"], center=kw["center"])
    assert rel_err(y, g["wav"]) < 1e-5


def test_num_frames_int_rule():
    lens = th.tensor([64000, 16000, 513])
    assert O.num_frames(lens, 512, 160, False).tolist() == [397, 97, 1]
    assert O.num_frames(lens, 400, 160, False).tolist() == [398, 98, 1]
    assert O.num_frames(lens, 512, 256, True).tolist() == [251, 63, 3]


@pytest.mark.reference
def test_oracle_vs_live_reference_features():
    import_reference()
    from aps.transform import AsrTransform
    th.manual_seed(0)
    x = 0.1 * th.randn(3, 20000)
    lens = th.tensor([20000, 15000, 9000])
    for mode, feats, power, lb, pb, an, center in itertools.product(
            ["librosa", "kaldi", "torch"], ["fbank-log-cmvn", "spectrogram-log", "emph-fbank-log-cmvn"],
            [False, True], [0, 1.0], [True, False], [True, False], [False, True]):
        kw = dict(feats=feats, stft_mode=mode, use_power=power, log_lower_bound=lb, norm_per_band=pb,
                  audio_norm=an, center=center)
        a, na = AsrTransform(**kw)(x.clone(), lens.clone())
        b, nb = O.AsrFeatures(O.AsrFeatCfg(**kw))(x.clone(), lens.clone())
        assert th.equal(na, nb), kw
        assert rel_err(b, a) < 1e-6, kw


@pytest.mark.reference
def test_mel_restatement_matches_shim_used_for_goldens():
    """oracle.mel_filterbank and the ref_shims librosa stand-in restate the same published formula."""
    import_reference()
    from aps.transform.utils import mel_filter
    for kw in (dict(frame_len=400), dict(frame_len=512, num_mels=40, fmin=50, fmax=-200), dict(frame_len=200, norm=True)):
        assert rel_err(O.mel_filterbank(**kw), mel_filter(**kw)) < 1e-6
