"""The reference's OWN test fixtures: tests/python/test_transform.py:102-150 runs AsrTransform on
tests/data/transform/egs1.wav (807 frames; dims 257 / 80 / 13 / 39) and EnhTransform on the 5-channel egs2.wav
(packed STFT [1, 5, 257, 366, 2], IPD features [1, 366, 257*4]).  The committed goldens hold the int16 samples and
the values the live reference computes for them (every 8th frame; oracle/gen_golden.py --only-fixtures)."""
import json

import numpy as np
import pytest
import torch as th

from conftest import FLOAT_TOL, load_golden, rel_err
from oracle import transform as O

DEV = "cuda:0"
CASES = [(m, f) for m in ("librosa", "torch")
         for f in ("spectrogram-log", "emph-fbank-log-cmvn", "mfcc", "mfcc-splice", "mfcc-delta")]
SHAPES = {"spectrogram-log": [1, 807, 257], "emph-fbank-log-cmvn": [1, 807, 80], "mfcc": [1, 807, 13],
          "mfcc-splice": [1, 807, 39], "mfcc-delta": [1, 807, 39]}        # what the reference's tests assert


def _egs1():
    cfg, g = load_golden("ref_egs1")
    return cfg, g, (g["pcm"].float() / 32768.0)[None]


@pytest.mark.parametrize("mode,feats", CASES)
def test_oracle_on_reference_fixture_egs1(mode, feats):
    cfg, g, wav = _egs1()
    y, _ = O.AsrFeatures(O.AsrFeatCfg(feats=feats, stft_mode=mode, frame_len=400, frame_hop=160, use_power=True,
                                      pre_emphasis=0.96))(wav.clone(), None)
    assert list(y.shape) == SHAPES[feats] == cfg["shapes"][f"{mode}.{feats}"]
    assert int(th.isnan(y).sum()) == 0
    assert rel_err(y[:, ::cfg["step"]], g[f"{mode}.{feats}"]) < 2e-5


def test_oracle_on_reference_fixture_egs2():
    cfg, g = load_golden("ref_egs2")
    wav = (g["pcm"].float() / 32768.0).t()[None]                            # 1 x 5 x S
    K, w = O.dft_kernel(512, O.window("sqrthann", 512))
    packed = O.stft_dense(wav, K, w, 256)
    assert list(packed.shape) == cfg["packed_shape"] == [1, 5, 257, 366, 2]
    assert rel_err(packed[..., ::cfg["step"], :], g["packed"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("mode,feats", CASES)
def test_asr_transform_on_reference_fixture_egs1(mode, feats):
    from aps_b200.transform import AsrTransform
    cfg, g, wav = _egs1()
    t = AsrTransform(feats=feats, stft_mode=mode, frame_len=400, frame_hop=160, use_power=True, pre_emphasis=0.96).to(DEV)
    y, _ = t(wav.to(DEV), None)
    assert list(y.shape) == SHAPES[feats] and t.feats_dim == SHAPES[feats][-1]
    assert int(th.isnan(y).sum()) == 0
    assert rel_err(y[:, ::cfg["step"]], g[f"{mode}.{feats}"]) < FLOAT_TOL


@pytest.mark.gpu
def test_enh_transform_on_reference_fixture_egs2():
    from aps_b200.transform import EnhTransform
    cfg, g = load_golden("ref_egs2")
    wav = (g["pcm"].float() / 32768.0).t()[None].contiguous()
    t = EnhTransform(feats="ipd", frame_len=512, frame_hop=256, ipd_index="0,1;0,2;0,3;0,4").to(DEV)
    packed, _ = t.encode(wav.to(DEV), None)
    feats = t(packed)
    assert list(packed.shape) == [1, 5, 257, 366, 2] and list(feats.shape) == [1, 366, 257 * 4]
    assert t.feats_dim == 257 * 4 and int(th.isnan(feats).sum()) == 0
    assert rel_err(packed[..., ::cfg["step"], :], g["packed"]) < FLOAT_TOL
    # IPD = cos of a phase difference: bins whose magnitude is at the noise floor have an ill-conditioned phase, so the
    # comparison is on the energy-weighted features (the reference's own test only checks shape / NaN here)
    ref = g["feats"]
    mag = O.magnitude(g["packed"][0]).permute(2, 1, 0)                       # T' x F x C
    wgt = (mag[..., 0] > 1e-3 * mag.max()).float().repeat(1, 4)[None]
    assert float(((feats[:, ::cfg["step"]].cpu() - ref).abs() * wgt).max()) < 5e-3
