"""Generate tests/golden/*.npz from the LIVE reference (funcwj/aps at /root/reference).

Run in the build container only (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py

Every file stores the seeded INPUTS next to the reference's OUTPUTS and the constructor kwargs
(as a JSON string), so the tests need neither the reference nor an RNG-compatible torch build.
The reference is imported unmodified; `oracle/ref_shims/` only supplies the three third-party
modules that are absent from this image (librosa.filters.mel, kaldi_python_io, soundfile).
"""
import json
import os
import sys
import warnings

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("APS_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims"), REF]
warnings.filterwarnings("ignore")
OUT = os.path.join(ROOT, "tests", "golden")


def save(name, kwargs, **arrays):
    arrays = {k: (v.detach().cpu().numpy() if isinstance(v, th.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kwargs=json.dumps(kwargs), **arrays)
    print(f"{name}: " + ", ".join(f"{k}{list(v.shape)}" for k, v in arrays.items()))


def wave(seed, *shape, kind="randn"):
    g = th.Generator().manual_seed(seed)
    if kind == "randn":
        return 0.1 * th.randn(*shape, generator=g)
    return th.rand(*shape, generator=g)


def objf_cases():
    """Si-SNR / SNR / PIT values of aps/task/objf.py on seeded estimate/reference pairs (row a25)."""
    from aps.task.objf import hybrid_permu_objf, permu_invarint_objf, sisnr_objf, snr_objf
    th.set_num_threads(4)
    g = th.Generator().manual_seed(2500)
    N, S = 6, 4000
    refs = [0.1 * th.randn(N, S, generator=g) + 0.02 * k for k in range(3)]
    # estimates: scaled references + noise at graded levels (about 40 dB ... -5 dB), speakers swapped for odd n
    level = th.tensor([0.001, 0.003, 0.01, 0.03, 0.1, 0.2])[:, None]
    ests = [(0.5 + 0.3 * k) * refs[k] + level * th.randn(N, S, generator=g) + 0.01 for k in range(3)]
    swap = th.arange(N) % 2 == 1
    e0, e1 = ests[0].clone(), ests[1].clone()
    e0[swap], e1[swap] = ests[1][swap], ests[0][swap]
    ests = [e0, e1, ests[2]]
    arrays = {f"ref{k}": refs[k] for k in range(3)}
    arrays.update({f"est{k}": ests[k] for k in range(3)})
    for zm in (True, False):
        for nn_ in (True, False):
            arrays[f"sisnr_zm{int(zm)}_nn{int(nn_)}"] = sisnr_objf(ests[0], refs[0], zero_mean=zm, non_nagetive=nn_)
    arrays["snr"] = snr_objf(ests[0], refs[0])
    arrays["snr_nn"] = snr_objf(ests[0], refs[0], non_nagetive=True)
    arrays["snr_max30"] = snr_objf(ests[0], refs[0], snr_max=30)
    neg = lambda x, s: -sisnr_objf(x, s)
    for K in (2, 3):
        loss, index = permu_invarint_objf(ests[:K], refs[:K], neg, return_permutation=True)
        arrays[f"pit{K}_loss"], arrays[f"pit{K}_index"] = loss, index
    arrays["hybrid_3of2"] = hybrid_permu_objf(ests, refs, neg, permute=True, permu_num_spks=2)
    arrays["hybrid_nopermute"] = hybrid_permu_objf(ests, refs, neg, weight=[0.5, 0.3, 0.2], permute=False)
    save("objf_0", dict(N=N, S=S), **arrays)


def freqsa_cases():
    """Frequency-domain spectral-approximation tasks (aps/task/sse.py:207-455, row f1): the live reference's
    LinearFreqSaTask / MelFreqSaTask on seeded mixtures with a stub network that returns fixed masks."""
    import torch.nn as nn
    from aps.task.sse import LinearFreqSaTask, MelFreqSaTask
    from aps.transform import EnhTransform
    th.set_num_threads(4)
    g = th.Generator().manual_seed(2600)
    N, S = 4, 4000
    enh_kw = dict(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann", center=True)
    enh = EnhTransform(**enh_kw)
    refs = [0.1 * th.randn(N, S, generator=g) for _ in range(2)]
    mix = refs[0] + refs[1] + 0.01 * th.randn(N, S, generator=g)
    T = int(enh.num_frames(th.tensor([S]))[0])
    masks = [th.rand(N, 257, T, generator=g) for _ in range(2)]
    masks[0][1::2], masks[1][1::2] = masks[1][1::2].clone(), masks[0][1::2].clone()

    class Stub(nn.Module):
        def __init__(self):
            super().__init__()
            self.enh_transform = enh

        def forward(self, mix):
            return masks

    egs = {"mix": mix, "ref": refs}
    arrays = {"mix": mix, "ref0": refs[0], "ref1": refs[1], "mask0": masks[0], "mask1": masks[1]}
    cfgs = {
        "lin_l2": ("linear", dict()),
        "lin_l1_psa": ("linear", dict(objf="L1", phase_sensitive=True)),
        "lin_tpsa_nopermute": ("linear", dict(phase_sensitive=True, truncated=1.0, permute=False, weight="0.7,0.3")),
        "lin_mapping": ("linear", dict(masking=False)),
        "mel_plain": ("mel", dict()),
        "mel_log_power": ("mel", dict(mel_log=True, power_mag=True, mel_scale=10, num_mels=40, phase_sensitive=True)),
    }
    with th.no_grad():
        for name, (kind, kw) in cfgs.items():
            task = (LinearFreqSaTask if kind == "linear" else MelFreqSaTask)(Stub(), **kw)
            arrays["loss_" + name] = task(egs)["loss"]
    save("freqsa_0", dict(N=N, S=S, enh=enh_kw, cfgs={k: [v[0], v[1]] for k, v in cfgs.items()}), **arrays)


def timesa_cases():
    """Time-domain spectral-approximation and complex mapping / masking tasks (aps/task/sse.py:458-841, row f1) of the
    live reference, with stub networks that return fixed waveforms / spectra / masks."""
    import torch.nn as nn
    from aps.task.sse import ComplexMappingTask, ComplexMaskingTask, LinearTimeSaTask, MelTimeSaTask
    from aps.transform import EnhTransform
    th.set_num_threads(4)
    g = th.Generator().manual_seed(2700)
    N, S = 4, 4000
    enh_kw = dict(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann", center=True)
    enh = EnhTransform(**enh_kw)
    refs = [0.1 * th.randn(N, S, generator=g) for _ in range(2)]
    mix = refs[0] + refs[1] + 0.01 * th.randn(N, S, generator=g)
    ests = [refs[k] + 0.03 * th.randn(N, S, generator=g) for k in range(2)]
    ests[0][1::2], ests[1][1::2] = ests[1][1::2].clone(), ests[0][1::2].clone()
    T = int(enh.num_frames(th.tensor([S]))[0])
    spec = [0.5 * th.randn(N, 257, T, 2, generator=g) for _ in range(2)]       # "network" spectra / complex masks

    class Stub(nn.Module):
        def __init__(self, out):
            super().__init__()
            self.enh_transform = enh
            self.out = out

        def forward(self, mix):
            return [o.clone() for o in self.out]                               # TimeSaTask pre-emphasises in place

    arrays = {"mix": mix, "ref0": refs[0], "ref1": refs[1], "est0": ests[0], "est1": ests[1], "spec0": spec[0],
              "spec1": spec[1]}
    cfgs = {
        "tlin_l2": ("time_linear", dict()),
        "tlin_l1_center": ("time_linear", dict(objf="L1", center=True, window="hann", frame_len=400, frame_hop=160)),
        "tmel_log": ("time_mel", dict(mel_log=True, mel_scale=5, num_mels=40, permute=False)),
        "cmap_l1_mag": ("complex_mapping", dict()),
        "cmap_l2": ("complex_mapping", dict(objf="L2", add_magnitude_loss=False)),
        "cmask": ("complex_masking", dict()),
    }
    table = {"time_linear": (LinearTimeSaTask, ests), "time_mel": (MelTimeSaTask, ests),
             "complex_mapping": (ComplexMappingTask, spec), "complex_masking": (ComplexMaskingTask, spec)}
    with th.no_grad():
        for name, (kind, kw) in cfgs.items():
            cls, out = table[kind]
            egs = {"mix": mix.clone(), "ref": [r.clone() for r in refs]}
            arrays["loss_" + name] = cls(Stub(out), **kw)(egs)["loss"]
        # pre-emphasis lives in TimeSaTask only (the registered subclasses do not expose it): set it on the instance
        from aps.task.sse import WaTask
        for objf in ("L1", "L2"):
            arrays["loss_wa_" + objf] = WaTask(Stub(ests), objf=objf)({"mix": mix, "ref": refs})["loss"]
        task = LinearTimeSaTask(Stub(ests))
        task.pre_emphasis = 0.97
        arrays["loss_tlin_preemph"] = task({"mix": mix.clone(), "ref": [r.clone() for r in refs]})["loss"]
    save("timesa_0", dict(N=N, S=S, enh=enh_kw, cfgs={k: [v[0], v[1]] for k, v in cfgs.items()}), **arrays)


def norm_cases():
    """Per-utterance normalisations over time: TCN with cLN / gLN / IN (tcn.py:75-88) and the transformer
    encoder behind LinearProj(norm="LN") (proj.py:30-56, component.py:86-114)."""
    import copy
    from aps.asr.transformer.encoder import TransformerEncoder
    from aps.sse.bss.tcn import FreqConvTasNet
    from aps.transform import EnhTransform
    th.set_num_threads(4)
    for i, norm in enumerate(["cLN", "gLN", "IN"]):
        g = th.Generator().manual_seed(630 + i)
        ekw = dict(feats="spectrogram-log-cmvn", frame_len=128, frame_hop=64)
        nkw = dict(in_features=65, num_bins=65, B=2, N=2, K=3, conv_channels=24, proj_channels=16, norm=norm,
                   num_spks=2 if i == 0 else 1, non_linear="sigmoid" if i else "relu")
        net = FreqConvTasNet(enh_transform=EnhTransform(**ekw), **nkw).eval()
        with th.no_grad():
            for name, prm in net.named_parameters():
                if not name.startswith("enh_transform") and (prm.dim() <= 1 or name.endswith(("gamma", "beta"))):
                    prm.add_(0.1 * th.randn(prm.shape, generator=g))
        mix = wave(630 + i, 3, 2500)
        with th.no_grad():
            stft, _ = net.enh_transform.encode(mix, None)
            feats = net.enh_transform(stft)
            masks = net.mask_predict(feats)
            net.training_mode = "time"
            wav = net(mix)
        wav = th.stack(wav) if isinstance(wav, list) else wav
        arrays = dict(mix=mix, feats=feats, masks=masks, wav=wav)
        arrays.update({"p." + k: v for k, v in net.state_dict().items() if not k.endswith(".K")})
        save(f"tcn_{3 + i}", dict(enh=ekw, net=nkw), **arrays)
    for i, (arch, norm) in enumerate([("xfmr", "LN"), ("cfmr", "BN")]):
        g = th.Generator().manual_seed(560 + i)
        ak = dict(att_dim=64, nhead=2, feedforward_dim=96, att_dropout=0.1, ffn_dropout=0.1, pre_norm=bool(i))
        if arch == "cfmr":
            ak["kernel_size"] = 15
        cfg = dict(arch=arch, input_size=40, output_proj=-1, num_layers=2, proj="linear", proj_kwargs=dict(norm=norm),
                   pose="abs" if i == 0 else "rel", pose_kwargs=dict() if i == 0 else dict(lradius=30, rradius=30),
                   arch_kwargs=ak)
        net = TransformerEncoder(**copy.deepcopy(cfg)).eval()
        with th.no_grad():
            for name, buf in net.named_buffers():
                if name.endswith("running_mean"):
                    buf.copy_(0.2 * th.randn(buf.shape, generator=g))
                if name.endswith("running_var"):
                    buf.copy_(0.5 + th.rand(buf.shape, generator=g))
            for name, prm in net.named_parameters():
                if name.endswith("bias") or "norm" in name:
                    prm.add_(0.1 * th.randn(prm.shape, generator=g))
        T = 53
        x = th.randn(3, T, 40, generator=g)
        lens = th.tensor([T, T - 8, T - 23])
        with th.no_grad():
            y, yl = net(x, lens.clone())
        arrays = dict(x=x, lens=lens, y=y, ylens=yl)
        arrays.update({"p." + k: v for k, v in net.state_dict().items()})
        save(f"enc_{7 + i}", cfg, **arrays)


def time_tcn_cases():
    """Time-domain Conv-TasNet (aps/sse/bss/tcn.py:228-358) on seeded mixtures."""
    from aps.sse.bss.tcn import TimeConvTasNet
    th.set_num_threads(4)
    for i, kw in enumerate([dict(num_spks=2, non_linear="relu", norm="BN"),
                            dict(num_spks=2, non_linear="softmax", norm="cLN", skip_residual=True, scaling_param=True),
                            dict(num_spks=1, non_linear="sigmoid", norm="gLN", causal=True)]):
        g = th.Generator().manual_seed(660 + i)
        nkw = dict(L=20 if i != 1 else 16, N=48, X=3, R=2, B=16, H=24, P=3, **kw)
        net = TimeConvTasNet(**nkw).eval()
        with th.no_grad():
            for name, buf in net.named_buffers():
                if name.endswith("running_mean"):
                    buf.copy_(0.2 * th.randn(buf.shape, generator=g))
                if name.endswith("running_var"):
                    buf.copy_(0.5 + th.rand(buf.shape, generator=g))
            for name, prm in net.named_parameters():
                if prm.dim() <= 1 or name.endswith(("gamma", "beta")):
                    prm.add_(0.1 * th.randn(prm.shape, generator=g))
        mix = wave(660 + i, 3, 3003)
        with th.no_grad():
            wav = net(mix)
            one = net.infer(mix[1])
        stack = lambda v: th.stack(v) if isinstance(v, list) else v
        arrays = dict(mix=mix, wav=stack(wav), one=stack(one))
        arrays.update({"p." + k: v for k, v in net.state_dict().items()})
        save(f"timetcn_{i}", dict(net=nkw), **arrays)


def reference_fixture_cases():
    """The reference's OWN test fixtures (tests/python/test_transform.py:102-150 on tests/data/transform/egs1.wav and
    egs2.wav): the shapes its tests assert (807 frames; [1, 5, 257, 366, 2]) plus the values it computes.  Outputs are
    stored for every `STEP`-th frame to keep the files small; the int16 samples are stored losslessly."""
    from scipy.io import wavfile

    from aps.transform import AsrTransform, EnhTransform
    th.set_num_threads(4)
    STEP = 8
    data = os.path.join(REF, "tests", "data", "transform")
    _, w1 = wavfile.read(os.path.join(data, "egs1.wav"))
    x1 = th.from_numpy(w1.astype(np.float32) / 32768.0)[None]
    arrays, shapes = {"pcm": w1}, {}
    for mode in ("librosa", "torch"):
        for feats in ("spectrogram-log", "emph-fbank-log-cmvn", "mfcc", "mfcc-splice", "mfcc-delta"):
            t = AsrTransform(feats=feats, stft_mode=mode, frame_len=400, frame_hop=160, use_power=True, pre_emphasis=0.96)
            y, _ = t(x1.clone(), None)
            key = f"{mode}.{feats}"
            shapes[key] = list(y.shape)
            arrays[key] = y[:, ::STEP]
    save("ref_egs1", dict(step=STEP, shapes=shapes, frame_len=400, frame_hop=160, use_power=True, pre_emphasis=0.96), **arrays)
    _, w2 = wavfile.read(os.path.join(data, "egs2.wav"))
    x2 = th.from_numpy(w2.T.astype(np.float32) / 32768.0)[None]                # 1 x 5 x S
    t = EnhTransform(feats="ipd", frame_len=512, frame_hop=256, ipd_index="0,1;0,2;0,3;0,4")
    packed, _ = t.encode(x2, None)
    feats = t(packed)
    save("ref_egs2", dict(step=STEP, packed_shape=list(packed.shape), feats_shape=list(feats.shape), frame_len=512,
                          frame_hop=256, ipd_index="0,1;0,2;0,3;0,4"),
         pcm=w2, packed=packed[..., ::STEP, :], feats=feats[:, ::STEP])


def ctc_cases():
    """`asr@ctc` (CtcASR: AsrTransform -> conformer encoder with the CTC projection) through the reference's own
    factories: the golden of tests/test_dropin.py (the `_training_prep` chain of aps/asr/ctc.py:113-134)."""
    import copy
    from aps.libs import aps_asr_nnet, aps_transform
    tkw = dict(feats="fbank-log-cmvn", frame_len=400, frame_hop=160, window="hamm", pre_emphasis=0.97, num_mels=80)
    nkw = dict(input_size=80, vocab_size=40, ctc=True, ead=False, enc_type="cfmr",
               enc_kwargs=dict(arch_kwargs=dict(att_dim=128, nhead=2, feedforward_dim=1024, att_dropout=0.1, ffn_dropout=0.1,
                                                kernel_size=15), num_layers=1, proj="conv2d",
                               proj_kwargs=dict(conv_channels=32, num_layers=2), pose="rel",
                               pose_kwargs=dict(lradius=64, rradius=64)))
    g = th.Generator().manual_seed(900)
    th.manual_seed(900)
    net = aps_asr_nnet("asr@ctc")(asr_transform=aps_transform("asr")(**tkw), **copy.deepcopy(nkw)).eval()
    with th.no_grad():
        for name, buf in net.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.2 * th.randn(buf.shape, generator=g))
            if name.endswith("running_var"):
                buf.copy_(0.5 + th.rand(buf.shape, generator=g))
        for name, prm in net.named_parameters():
            if name.endswith("bias") or "norm" in name:
                prm.add_(0.1 * th.randn(prm.shape, generator=g))
    x = 0.1 * th.randn(6, 16000, generator=g)
    lens = th.tensor([16000, 16000, 14000, 12000, 9000, 6000])
    with th.no_grad():
        enc_out, enc_ctc, enc_len = net(x.clone(), lens.clone())
    arrays = dict(x=x, lens=lens, enc_out=enc_out, enc_len=enc_len)
    # the transform's buffers (DFT kernel, window, mel filters) are rebuilt by its constructor: only the encoder is stored
    arrays.update({"p." + k: v for k, v in net.encoder.state_dict().items()})
    save("ctc_0", dict(transform=tkw, net=nkw), **arrays)


def main():
    if "--only-ctc" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        return ctc_cases()
    if "--only-fixtures" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        return reference_fixture_cases()
    if "--only-timetcn" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        return time_tcn_cases()
    if "--only-norms" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        return norm_cases()
    if "--only-objf" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        return objf_cases()
    if "--only-freqsa" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        freqsa_cases()
        return timesa_cases()
    from aps.transform import AsrTransform, EnhTransform
    from aps.transform.utils import STFT, iSTFT
    th.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)

    # ---- AsrTransform: BASELINE config[0] (1 utt x 4 s) in the three stft modes ---------------------
    for mode in ("librosa", "kaldi", "torch"):
        kw = dict(feats="fbank-log-cmvn", frame_len=400, frame_hop=160, window="hamm", pre_emphasis=0.97,
                  num_mels=80, stft_mode=mode)
        x = wave(1, 1, 64000)
        y, n = AsrTransform(**kw)(x.clone(), th.tensor([64000]))
        if mode == "librosa":
            save("asr_c1_input", {}, wav=x)                      # shared by the three modes
        save(f"asr_c1_{mode}", kw, feats=y, num_frames=n)
    # ---- ragged batch, assorted epilogues ------------------------------------------------------------
    grid = [
        dict(feats="fbank-log-cmvn", stft_mode="kaldi", audio_norm=False, log_lower_bound=1.0),   # aishell 1e
        dict(feats="spectrogram-log-cmvn", stft_mode="librosa", use_power=True, norm_per_band=False),
        dict(feats="emph-fbank-log-cmvn", stft_mode="librosa", use_power=True, pre_emphasis=0.96),
        dict(feats="fbank-log", stft_mode="librosa", center=True, num_mels=40, window="hann"),
        dict(feats="spectrogram", stft_mode="torch", center=True, stft_normalized=True),
        dict(feats="fbank-log-cmvn", stft_mode="librosa", frame_len=200, frame_hop=80, norm_var=False,
             pre_emphasis=0.0),
        dict(feats="fbank-log-cmvn-splice-delta", stft_mode="librosa", lctx=1, rctx=1),
        dict(feats="mfcc", stft_mode="kaldi", lifter=22),
    ]
    for i, kw in enumerate(grid):
        x = wave(10 + i, 3, 8000, kind="rand" if i % 2 else "randn")
        lens = th.tensor([8000, 6500, 4000])
        y, n = AsrTransform(**kw)(x.clone(), lens.clone())
        save(f"asr_grid_{i}", kw, wav=x, lens=lens, feats=y, num_frames=n)
    # ---- STFT / iSTFT (the reference's own test sizes, tests/python/test_transform.py:21-37) ---------
    k = 0
    for mode in ("librosa", "kaldi", "torch"):
        for (fl, fh) in ((512, 256), (1024, 256), (256, 128), (400, 160)):
            for window, center in (("sqrthann", True), ("hamm", False)):
                if mode == "kaldi" and fl != 400:
                    continue
                kw = dict(frame_len=fl, frame_hop=fh, window=window, center=center, mode=mode)
                x = wave(100 + k, 2, 3, 3000) if k % 2 else wave(100 + k, 2, 3000)
                spec = STFT(**kw)(x)
                save(f"stft_{k}", kw, wav=x, spec=spec)
                if x.dim() == 2 and not (mode == "torch" and not center):
                    rec = iSTFT(**kw)(spec)
                    save(f"istft_{k}", kw, spec=spec, wav=rec)
                k += 1
    # polar in / out
    kw = dict(frame_len=512, frame_hop=256, window="sqrthann", center=False, mode="librosa")
    x = wave(200, 2, 4000)
    pol = STFT(**kw)(x, return_polar=True)
    save("stft_polar", kw, wav=x, spec=pol, rec=iSTFT(**kw)(pol, return_polar=True))
    # ---- EnhTransform: encode / forward(+IPD) / decode ---------------------------------------------------
    for i, kw in enumerate([
            dict(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann"),
            dict(feats="spectrogram-log-cmvn-ipd", frame_len=512, frame_hop=256, ipd_index="0,1;0,2;0,3",
                 cos_ipd=True, sin_ipd=True, ref_channel=1),
            dict(feats="fbank-log-ipd", frame_len=400, frame_hop=160, ipd_index="1,0", num_mels=40, center=True),
    ]):
        x = wave(300 + i, 2, 4, 4000)
        t = EnhTransform(**kw)
        packed, n = t.encode(x, th.tensor([4000, 3500]))
        feats = t(packed)
        rec = t.decode([packed[:, 0]])[0]
        save(f"enh_{i}", kw, wav=x, packed=packed, num_frames=n, feats=feats, rec=rec)
    # ---- MVDR front-end (aps/asr/filter/mvdr.py): covariance, reference attention, weights, beamforming -----
    from aps.asr.filter.mvdr import MvdrBeamformer, estimate_covar
    from aps.cplx import ComplexTensor
    for i, (N, C, Fb, T, use_n, use_len, norm) in enumerate([(2, 4, 65, 30, False, True, True),
                                                             (3, 5, 33, 47, True, True, True),
                                                             (2, 2, 129, 20, True, False, False)]):
        g = th.Generator().manual_seed(400 + i)
        net = MvdrBeamformer(Fb, att_dim=24, mask_norm=norm).eval()
        with th.no_grad():
            for prm in net.parameters():
                prm.copy_(th.randn(prm.shape, generator=g) * 0.3)
        xr, xi = th.randn(N, C, Fb, T, generator=g), th.randn(N, C, Fb, T, generator=g)
        ms, mn = th.rand(N, T, Fb, generator=g), th.rand(N, T, Fb, generator=g)
        lens = th.tensor([T, T - 7, T - 11][:N])
        with th.no_grad():
            y = net(ms, ComplexTensor(xr, xi), mask_n=mn if use_n else None, x_len=lens if use_len else None)
            R = estimate_covar(ms.transpose(1, 2), ComplexTensor(xr, xi))
        arrays = dict(xr=xr, xi=xi, mask_s=ms, mask_n=mn, lens=lens, yr=y.real, yi=y.imag, Rr=R.real, Ri=R.imag)
        arrays.update({"p." + k: v for k, v in net.state_dict().items()})
        save(f"mvdr_{i}", dict(num_bins=Fb, att_dim=24, mask_norm=norm, use_n=use_n, use_len=use_len), **arrays)
    # ---- transformer / conformer encoders (aps/asr/transformer/encoder.py) -------------------------------------
    import copy
    from aps.asr.transformer.encoder import TransformerEncoder
    ak_c = dict(att_dim=64, nhead=2, feedforward_dim=96, att_dropout=0.1, ffn_dropout=0.1, kernel_size=15)
    ak_x = dict(att_dim=64, nhead=2, feedforward_dim=96, att_dropout=0.1, ffn_dropout=0.1)
    variants = [
        ("cfmr", "rel", False, dict(conv_channels=16, num_layers=3), dict(lradius=12, rradius=9), 77, -1),
        ("cfmr", "rel", True, dict(conv_channels=16, num_layers=2), dict(lradius=40, rradius=40), 61, 20),
        ("cfmr", "abs", False, dict(conv_channels=16, num_layers=2), dict(), 61, -1),
        ("cfmr", "xl", False, dict(conv_channels=16, num_layers=2), dict(), 61, -1),
        ("xfmr", "abs", False, dict(conv_channels=16, num_layers=2), dict(scaled=True), 61, -1),
        ("xfmr", "rel", True, dict(conv_channels=16, num_layers=2), dict(lradius=40, rradius=40), 61, -1),
        ("xfmr", "xl", True, dict(conv_channels=16, num_layers=2), dict(), 61, -1),
    ]
    for i, (arch, pose, pre, pkw, posekw, T, outp) in enumerate(variants):
        g = th.Generator().manual_seed(500 + i)
        cfg = dict(arch=arch, input_size=40, output_proj=outp, num_layers=2, proj="conv2d", proj_kwargs=pkw,
                   pose=pose, pose_kwargs=posekw, arch_kwargs=dict(ak_c if arch == "cfmr" else ak_x, pre_norm=pre))
        net = TransformerEncoder(**copy.deepcopy(cfg)).eval()
        with th.no_grad():                                  # non-trivial BatchNorm statistics and biases
            for name, buf in net.named_buffers():
                if name.endswith("running_mean"):
                    buf.copy_(0.2 * th.randn(buf.shape, generator=g))
                if name.endswith("running_var"):
                    buf.copy_(0.5 + th.rand(buf.shape, generator=g))
            for name, prm in net.named_parameters():
                if name.endswith("bias") or "norm" in name:
                    prm.add_(0.1 * th.randn(prm.shape, generator=g))
        x = th.randn(3, T, 40, generator=g)
        lens = th.tensor([T, T - 8, T - 23])
        with th.no_grad():
            y, yl = net(x, lens.clone())
        arrays = dict(x=x, lens=lens, y=y, ylens=yl)
        arrays.update({"p." + k: v for k, v in net.state_dict().items()})
        save(f"enc_{i}", cfg, **arrays)
    # ---- frequency-domain Conv-TasNet mask estimator (aps/sse/bss/tcn.py) -----------------------------------------
    from aps.sse.bss.tcn import FreqConvTasNet
    for i, kw in enumerate([dict(num_spks=2, non_linear="relu"),
                            dict(num_spks=1, non_linear="sigmoid", skip_residual=True, scaling_param=True),
                            dict(num_spks=1, non_linear="sigmoid", causal=True)]):
        g = th.Generator().manual_seed(600 + i)
        ekw = dict(feats="spectrogram-log-cmvn", frame_len=128, frame_hop=64)
        nkw = dict(in_features=65, num_bins=65, B=3, N=3 if i == 1 else 2, K=3, conv_channels=24, proj_channels=16, **kw)
        net = FreqConvTasNet(enh_transform=EnhTransform(**ekw), **nkw).eval()
        with th.no_grad():
            for name, buf in net.named_buffers():
                if name.endswith("running_mean"):
                    buf.copy_(0.2 * th.randn(buf.shape, generator=g))
                if name.endswith("running_var"):
                    buf.copy_(0.5 + th.rand(buf.shape, generator=g))
            for name, prm in net.named_parameters():
                if not name.startswith("enh_transform") and prm.dim() <= 1:
                    prm.add_(0.1 * th.randn(prm.shape, generator=g))
        mix = wave(600 + i, 2, 2500)
        with th.no_grad():
            stft, _ = net.enh_transform.encode(mix, None)
            feats = net.enh_transform(stft)
            masks = net.mask_predict(feats)
            net.training_mode = "time"
            wav = net(mix)
        wav = th.stack(wav) if isinstance(wav, list) else wav
        arrays = dict(mix=mix, feats=feats, masks=masks, wav=wav)
        arrays.update({"p." + k: v for k, v in net.state_dict().items() if not k.endswith(".K")})
        save(f"tcn_{i}", dict(enh=ekw, net=nkw), **arrays)
    # ---- DCCRN (aps/sse/bss/dccrn.py), the configuration pinned by the reference's tests -------------------------
    from aps.sse.bss.dccrn import DCCRN
    for i, (spk, conn, nl, C) in enumerate([(1, "cat", "sigmoid", "4,8,8,8,16,16,32"), (2, "cat", "tanh", "4,8,8,8,16,16,32")]):
        g = th.Generator().manual_seed(700 + i)
        ekw = dict(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, center=True)
        nkw = dict(cplx=True, K="3,3;3,3;3,3;3,3;3,3;3,3;3,3", S="2,1;2,1;2,1;2,1;2,1;2,1;2,1", P="1,1,1,1,1,0,0",
                   O="0,0,0,0,0,0,1", C=C, num_spks=spk, rnn_resize=64, rnn_hidden=40, non_linear=nl, connection=conn)
        net = DCCRN(enh_transform=EnhTransform(**ekw), **nkw).eval()
        with th.no_grad():
            for name, buf in net.named_buffers():
                if name.endswith("running_mean"):
                    buf.copy_(0.2 * th.randn(buf.shape, generator=g))
                if name.endswith("running_var"):
                    buf.copy_(0.5 + th.rand(buf.shape, generator=g))
        mix = wave(700 + i, 2, 6000, kind="rand")
        with th.no_grad():
            wav = net(mix)
            net.training_mode = "freq"
            msk = net(mix)
        stack = lambda v: th.stack(v) if isinstance(v, list) else v
        arrays = dict(mix=mix, wav=stack(wav), masks=stack(msk))
        arrays.update({"p." + k: v for k, v in net.state_dict().items() if not k.endswith(".K")})  # K: 2 MB each
        save(f"dccrn_{i}", dict(enh=ekw, net=nkw), **arrays)
    objf_cases()
    freqsa_cases()
    timesa_cases()
    norm_cases()
    time_tcn_cases()
    reference_fixture_cases()
    # state-dict layout of the recipe transform (conf/asr/aishell_v1/1e.yaml:17-40)
    t = AsrTransform(feats="perturb-fbank-log-cmvn-aug", frame_len=400, frame_hop=160, window="hamm",
                     audio_norm=False, pre_emphasis=0.97, stft_mode="kaldi", log_lower_bound=1, num_mels=80)
    sd = {k: list(v.shape) for k, v in t.state_dict().items()}
    e = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256)
    sd_e = {k: list(v.shape) for k, v in e.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_layout.json"), "w") as fd:
        json.dump({"asr_aishell_1e": sd, "enh_default": sd_e}, fd, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
