"""GPU parity of the remaining feature-chain tokens: the kernels of csrc/featops.cu
behind SpecAugTransform / SpliceTransform / DeltaTransform / SpeedPerturbTransform / DiscreteCosineTransform against the
same layers evaluated with plain tensor ops on the CPU — which tests/test_host_logic.py pins to the live reference layers
(asr.py:116-195, :467-517, :621-781) — and against the oracle restatements."""
import random

import pytest
import torch as th

from conftest import FLOAT_TOL, rel_err
from oracle import transform as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_splice_and_subsampling_are_bit_exact():
    import aps_b200.transform.asr as A
    th.manual_seed(0)
    for shape in ((2, 30, 40), (3, 2, 17, 13), (1, 5, 8)):
        x = th.randn(*shape)
        for (l, r, sub) in ((2, 1, 2), (0, 0, 3), (3, 3, 1), (0, 2, 1), (1, 0, 4)):
            lay = A.SpliceTransform(l, r, sub)
            got = lay(x.to(DEV))
            assert th.equal(got.cpu(), lay(x)), (shape, l, r, sub)
            ref = O.splice(x, l, r)
            if sub != 1:
                ref = ref[..., :(x.shape[-2] // sub) * sub:sub, :]
            assert th.equal(got.cpu(), ref)


def test_delta_features():
    import aps_b200.transform.asr as A
    th.manual_seed(1)
    x = th.randn(3, 41, 23)
    for ctx, order in ((2, 2), (1, 1), (3, 2)):
        lay = A.DeltaTransform(ctx, order)
        got = lay(x.to(DEV))
        assert got.shape == (3, 41, 23 * (order + 1))
        assert rel_err(got, lay(x)) < 1e-6
        assert rel_err(got, O.delta(x, ctx, order)) < 1e-6
        ch = A.DeltaTransform(ctx, order, delta_as_channel=True)
        assert rel_err(ch(x.to(DEV)), ch(x)) < 1e-6 and ch(x.to(DEV)).shape == (3, order + 1, 41, 23)
    x4 = th.randn(2, 3, 20, 8)
    lay = A.DeltaTransform(2, 2)
    assert rel_err(lay(x4.to(DEV)), lay(x4)) < 1e-6


def test_dct_and_unfused_mel_run_on_this_packages_gemm():
    import aps_b200.transform.asr as A
    from aps_b200 import _lib
    th.manual_seed(2)
    x = th.randn(4, 50, 40)
    lay = A.DiscreteCosineTransform(13, 40, 22)
    c0 = _lib.CALLS
    got = lay.to(DEV)(x.to(DEV))
    assert _lib.CALLS > c0, "the DCT must go through the C ABI (ops.linear), not a library matmul"
    assert rel_err(got, A.DiscreteCosineTransform(13, 40, 22)(x)) < 1e-5
    assert rel_err(got[..., :13], th.nn.functional.linear(x, O.dct_matrix(13, 40)) * lay.cepstral_lifter.cpu()) < 1e-5
    mel = A.MelTransform(400, num_mels=80)
    s = th.rand(2, 30, 257)
    assert rel_err(mel.to(DEV)(s.to(DEV)), A.MelTransform(400, num_mels=80)(s)) < 1e-5


@pytest.mark.parametrize("mask_zero", [True, False])
def test_specaug_apply_kernel(mask_zero):
    import aps_b200.transform.asr as A
    from aps_b200 import ops
    th.manual_seed(3)
    for shape in ((3, 50, 20), (2, 2, 33, 17)):
        x = th.randn(*shape) + 0.3
        N, T, F = shape[0], shape[-2], shape[-1]
        random.seed(5)
        mask = A.tf_mask(N, (T, F), max_bands=8, max_frame=12, num_freq_masks=2, num_time_masks=2)
        got = ops.specaug_apply(x.to(DEV), mask.to(DEV), mask_zero)
        m = mask if x.dim() == 3 else mask.unsqueeze(1)
        ref = x * m if mask_zero else th.masked_fill(x, m == 0, x.mean())
        assert rel_err(got, ref) < 1e-6
        assert rel_err(got, O.specaug_apply(x, mask, mask_zero)) < 1e-6
        if mask_zero:
            assert th.equal(got.cpu(), ref)


def test_specaug_mean_fill_through_the_transform():
    """fbank-log-cmvn-aug with aug_mask_zero=False: fused feature kernel, then the mean-fill kernel (global mean of the
    final features, asr.py:680-683)."""
    from aps_b200.transform import AsrTransform
    from aps_b200.transform.asr import tf_mask
    kw = dict(feats="fbank-log-cmvn-aug", frame_len=400, frame_hop=160, window="hamm", pre_emphasis=0.97, num_mels=80,
              aug_prob=1.0, aug_time_args=(10, 2), aug_freq_args=(8, 2), aug_mask_zero=False)
    t = AsrTransform(**kw).to(DEV).train()
    x = (0.1 * th.randn(3, 8000)).to(DEV)
    base, _ = AsrTransform(**dict(kw, feats="fbank-log-cmvn")).to(DEV).eval()(x, None)
    random.seed(11)
    got, _ = t(x, None)
    random.seed(11)
    mask = tf_mask(3, tuple(base.shape[1:]), max_bands=8, max_frame=10, num_freq_masks=2, num_time_masks=2, device=DEV)
    assert rel_err(got, th.masked_fill(base, mask == 0, base.mean())) < 1e-6


def test_speed_perturb_kernel_train_mode():
    """Train-mode SpeedPerturbTransform on the GPU (one launch for the batch) against the same layer on the CPU with the
    same RNG state: per-utterance choices, zero padding, output lengths."""
    import aps_b200.transform.asr as A
    lay = A.SpeedPerturbTransform(sr=16000, perturb="0.9,1.0,1.1").train()
    dev_lay = A.SpeedPerturbTransform(sr=16000, perturb="0.9,1.0,1.1").to(DEV).train()
    x = 0.1 * th.randn(7, 16000)
    lens = th.tensor([16000, 15000, 14000, 12000, 9000, 8000, 4000])
    for seed in (0, 1, 2):
        th.manual_seed(seed)
        ref = lay(x)
        th.manual_seed(seed)
        got = dev_lay(x.to(DEV))
        assert th.equal(lay.last_choice, dev_lay.last_choice)
        assert got.shape == ref.shape
        assert rel_err(got, ref) < 1e-5
        assert th.equal(lay.output_length(lens), dev_lay.output_length(lens))
    with pytest.raises(RuntimeError):
        dev_lay(th.randn(2, 5, device=DEV))
    # train-mode transform end to end: perturb -> fused fbank; frame counts follow the perturbed lengths
    from aps_b200.transform import AsrTransform
    t = AsrTransform(feats="perturb-fbank-log-cmvn", frame_len=400, frame_hop=160, window="hamm", pre_emphasis=0.97,
                     num_mels=80).to(DEV).train()
    th.manual_seed(4)
    feats, nf = t(x.to(DEV), lens.clone())
    assert feats.shape[0] == 7 and int(nf.max()) == feats.shape[1] and not th.isnan(feats).any()
