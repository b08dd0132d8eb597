"""CPU-side logic of the Python shells: parameter layout, constants, integer rules, error behaviour."""
import json
import os

import pytest
import torch as th

from conftest import GOLDEN, import_reference, rel_err
from oracle import transform as O


def test_state_dict_layout_matches_reference():
    from aps_b200.transform import AsrTransform, EnhTransform
    layout = json.load(open(os.path.join(GOLDEN, "state_dict_layout.json")))
    t = AsrTransform(feats="perturb-fbank-log-cmvn-aug", frame_len=400, frame_hop=160, window="hamm",
                     audio_norm=False, pre_emphasis=0.97, stft_mode="kaldi", log_lower_bound=1, num_mels=80)
    assert {k: list(v.shape) for k, v in t.state_dict().items()} == layout["asr_aishell_1e"]
    e = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256)
    assert {k: list(v.shape) for k, v in e.state_dict().items()} == layout["enh_default"]


@pytest.mark.parametrize("mode", ["librosa", "kaldi"])
@pytest.mark.parametrize("frame_len", [400, 512, 200])
@pytest.mark.parametrize("inverse,normalized", [(False, False), (True, False), (False, True)])
def test_kernel_and_window_values(mode, frame_len, inverse, normalized):
    from aps_b200.transform.utils import init_kernel, init_window
    for wnd in ("hamm", "sqrthann", "hann", "blackman", "bartlett", "rect"):
        assert th.equal(init_window(wnd, frame_len), O.window(wnd, frame_len))
    K, w = init_kernel(frame_len, 160, init_window("hamm", frame_len), normalized=normalized, inverse=inverse, mode=mode)
    Ko, wo = O.dft_kernel(frame_len, O.window("hamm", frame_len), True, normalized, inverse, mode)
    assert K.shape == Ko.shape and th.equal(w, wo)
    assert (K - Ko).abs().max() < 2e-6 * Ko.abs().max()


def test_mel_filter_values():
    from aps_b200.transform.utils import mel_filter
    for kw in (dict(frame_len=400), dict(frame_len=512, num_mels=40, fmin=50, fmax=-200), dict(frame_len=200, norm=True)):
        a, b = mel_filter(**kw), O.mel_filterbank(**kw)
        assert a.shape == b.shape and rel_err(a, b) < 1e-6
    assert int((mel_filter(400) != 0).sum()) == 503


def test_num_frames_quirks():
    from aps_b200.transform import AsrTransform, EnhTransform
    for mode, expect in (("librosa", 397), ("kaldi", 398), ("torch", 397)):
        t = AsrTransform(stft_mode=mode)
        assert t.num_frames(th.tensor([64000])).tolist() == [expect]
    e = EnhTransform(center=True)
    lens = th.tensor([64000])
    assert e.num_frames(lens).tolist() == [251]
    assert lens.tolist() == [64512]          # Q5: in-place side effect of the reference is kept
    with pytest.raises(AssertionError):
        AsrTransform().num_frames(th.tensor([512]))


def test_error_behaviour_without_gpu_work():
    from aps_b200.transform import AsrTransform
    from aps_b200.transform.utils import init_kernel, init_window
    with pytest.raises(RuntimeError, match="Unknown window"):
        init_window("kaiser", 400)
    with pytest.raises(ValueError, match="Unsupported mode"):
        init_kernel(400, 160, th.ones(400), mode="torch")
    with pytest.raises(RuntimeError, match="Unknown token"):
        AsrTransform(feats="fbank-foo")
    with pytest.raises(ValueError):
        AsrTransform(feats="")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        AsrTransform()(th.rand(1, 16000), None)              # the product path never falls back to the CPU


def test_specaug_masks_follow_reference_rng_order():
    import random

    from aps_b200.transform.asr import tf_mask
    random.seed(7)
    m = tf_mask(3, (50, 20), max_bands=8, max_frame=12, num_freq_masks=2, num_time_masks=2)
    assert m.shape == (3, 50, 20) and set(m.unique().tolist()) <= {0.0, 1.0}
    random.seed(7)
    assert th.equal(m, tf_mask(3, (50, 20), max_bands=8, max_frame=12, num_freq_masks=2, num_time_masks=2))


@pytest.mark.reference
def test_specaug_masks_equal_reference():
    import random
    import_reference()
    from aps.transform.augment import tf_mask as ref_mask

    from aps_b200.transform.asr import tf_mask
    for seed in range(3):
        random.seed(seed)
        a = ref_mask(4, (60, 30), pm=0.2, ps=0.2, max_bands=10, max_frame=15, num_freq_masks=2, num_time_masks=3)
        random.seed(seed)
        b = tf_mask(4, (60, 30), pm=0.2, ps=0.2, max_bands=10, max_frame=15, num_freq_masks=2, num_time_masks=3)
        assert th.equal(a, b)


@pytest.mark.reference
def test_layer_forwards_equal_reference_on_cpu():
    """The non-fused ("next" row) layers are plain tensor ops; check them against the reference layers."""
    import_reference()
    import aps.transform.asr as R

    import aps_b200.transform.asr as A
    th.manual_seed(0)
    x = th.randn(2, 30, 40)
    assert rel_err(A.DiscreteCosineTransform(13, 40, 22)(x), R.DiscreteCosineTransform(13, 40, 22)(x)) < 1e-6
    assert th.equal(A.SpliceTransform(2, 1, 2)(x), R.SpliceTransform(2, 1, 2)(x))
    assert rel_err(A.DeltaTransform(2, 2)(x), R.DeltaTransform(2, 2)(x)) < 1e-6
    a, r = A.SpeedPerturbTransform(), R.SpeedPerturbTransform()
    for wa, wr in zip(a.weights, r.weights):
        assert rel_err(wa, wr) < 1e-6
    assert th.equal(a.src_sr, r.src_sr) and th.equal(a.dst_sr, r.dst_sr)
    for cm in (dict(), dict(per_band=False), dict(norm_mean=False)):
        assert rel_err(A.CmvnTransform(**cm)(x), R.CmvnTransform(**cm)(x)) < 1e-6


def test_pack_guard_detects_in_place_updates():
    """derived weight copies (BN folding, TF32 splits, CUDA graphs) must be rebuilt after in-place parameter
    updates, not only after a top-level load_state_dict: the guard sees EMA-style updates, copy_ and a sub-module
    load_state_dict (all bump the version counter)."""
    import copy
    import time

    from aps_b200 import ops
    from aps_b200.asr.transformer import TransformerEncoder
    cfg = dict(arch="cfmr", input_size=80, num_layers=2, proj="conv2d", proj_kwargs=dict(conv_channels=32, num_layers=2),
               pose="rel", pose_kwargs=dict(lradius=8, rradius=8),
               arch_kwargs=dict(att_dim=64, nhead=4, feedforward_dim=128, kernel_size=7, pre_norm=False))
    net = TransformerEncoder(**copy.deepcopy(cfg)).eval()
    g = net._guard
    assert g.stale()                     # nothing built yet
    g.mark()
    assert not g.stale()
    with th.no_grad():
        net.encoder.layers[0].norm_ffn1.weight.mul_(0.999).add_(0.001)          # EMA-style update
    assert g.stale()
    g.mark()
    with th.no_grad():
        net.proj.conv.enc_layers[0].norm.norm.running_mean.copy_(th.ones(32))     # buffer update
    assert g.stale()
    g.mark()
    net.proj.load_state_dict(copy.deepcopy(net.proj.state_dict()))                # sub-module load
    assert g.stale()
    net._drop_packs()
    assert g.stale() and net._packs is None
    g.mark()
    t0 = time.perf_counter()
    for _ in range(100):
        g.stale()
    assert (time.perf_counter() - t0) / 100 < 2e-3      # cheap enough to run on every forward


def test_global_cmvn_stats_are_validated_before_the_kernel_sees_them():
    """gcmvn statistics keep their saved dtype (float64 from `th.load`) and may not match the feature width;
    the kernel must get a float32 copy of the right length or the reference's broadcast error — never a raw pointer."""
    import torch.nn as nn

    from aps_b200.transform.asr import CmvnTransform
    c = CmvnTransform(gcmvn="/nonexistent/cmvn.pt", dim=8)
    c.gmean = nn.Parameter(th.arange(8, dtype=th.float64), requires_grad=False)
    c.gstd = nn.Parameter(th.full((8,), 2.0, dtype=th.float64), requires_grad=False)
    dev = th.device("cpu")
    m, s = c.global_stats(dev, 8)
    assert m.dtype == th.float32 and s.dtype == th.float32 and m.is_contiguous()
    assert th.equal(m, th.arange(8, dtype=th.float32)) and th.equal(s, th.full((8,), 2.0))
    assert c.global_stats(dev, 8)[0] is m                       # cached
    with th.no_grad():
        c.gmean.add_(1.0)
    assert th.equal(c.global_stats(dev, 8)[0], th.arange(1, 9, dtype=th.float32))   # rebuilt after an in-place update
    with pytest.raises(RuntimeError, match="must match the size"):
        c.global_stats(dev, 80)
