"""Separation / enhancement tasks (network forward + objective value).

Time-domain Si-SNR / SNR on the fused objective kernel ("sse@sisnr", "sse@snr", aps/task/sse.py:60-167); the
frequency- and time-domain spectral-approximation tasks and the complex mapping / masking tasks
("sse@freq_linear_sa", "sse@freq_mel_sa", "sse@time_linear_sa", "sse@time_mel_sa", "sse@complex_mapping",
"sse@complex_masking", sse.py:207-841) on the fused STFT kernel plus device tensor ops.  Same constructor arguments
and `forward(egs) -> {"loss": scalar}` contract as the reference.  Forward (evaluation) only: the loss tensor carries
no autograd graph, training through it is out of scope.
"""
from typing import Dict, Optional

import torch as th
import torch.nn as nn

from .objf import FusedObjf, _Kind, hybrid_permu_objf


class Task(nn.Module):
    """aps/task/base.py:14-30"""

    def __init__(self, nnet: nn.Module, ctx: Optional[nn.Module] = None, description: str = "unknown") -> None:
        super(Task, self).__init__()
        self.nnet = nnet
        self.ctx = ctx
        self.description = description


def _mel_project(mel: th.Tensor, spec: th.Tensor) -> th.Tensor:
    """mel M x F applied to spec N x F x T -> N x M x T (aps/task/sse.py:430-440): on the device through this package's
    GEMM kernels (ops.project_rows), not a library matmul."""
    if spec.is_cuda:
        from .. import ops
        return ops.project_rows(spec.transpose(-1, -2), mel).transpose(-1, -2)
    return th.matmul(mel, spec)


class TimeDomainTask(Task):
    """aps/task/sse.py:26-103 (SepTask + TimeDomainTask)"""

    def __init__(self, nnet: nn.Module, num_spks: int = 2, permute: bool = True, description: str = "",
                 weight: Optional[str] = None) -> None:
        super(TimeDomainTask, self).__init__(nnet, description=description)
        self.weight = list(map(float, weight.split(","))) if weight is not None else None
        self.num_spks = num_spks
        self.permute = permute

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        raise NotImplementedError

    def fused_objf(self) -> FusedObjf:
        raise NotImplementedError

    def forward(self, egs: Dict) -> Dict:
        """egs: mix N x (C) x S, ref N x S or [N x S, ...]"""
        ref = egs["ref"]
        out = self.nnet(egs["mix"])
        if isinstance(out, th.Tensor):
            out, ref = [out], [ref]
        loss = hybrid_permu_objf(out, ref, self.fused_objf(), weight=self.weight, permute=self.permute,
                                 permu_num_spks=self.num_spks)
        return {"loss": th.mean(loss)}


class SisnrTask(TimeDomainTask):
    """Negative Si-SNR, optionally permutation invariant.  aps/task/sse.py:105-139"""

    def __init__(self, nnet: nn.Module, num_spks: int = 2, permute: bool = True, weight: Optional[str] = None,
                 zero_mean: bool = True, non_nagetive: bool = False) -> None:
        super(SisnrTask, self).__init__(nnet, num_spks=num_spks, permute=permute, weight=weight,
                                        description="Using SiSNR objective function for training")
        self.zero_mean = zero_mean
        self.non_nagetive = non_nagetive

    def fused_objf(self) -> FusedObjf:
        return FusedObjf(_Kind.SISNR, sign=-1.0, zero_mean=self.zero_mean, non_nagetive=self.non_nagetive)

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        return self.fused_objf()(out, ref)


class SnrTask(TimeDomainTask):
    """Negative SNR.  aps/task/sse.py:141-167"""

    def __init__(self, nnet: nn.Module, num_spks: int = 2, permute: bool = True, weight: Optional[str] = None,
                 snr_max: float = -1, non_nagetive: bool = False) -> None:
        super(SnrTask, self).__init__(nnet, num_spks=num_spks, permute=permute, weight=weight,
                                      description="Using SNR objective function for training")
        self.non_nagetive = non_nagetive
        self.snr_max = snr_max

    def fused_objf(self) -> FusedObjf:
        return FusedObjf(_Kind.SNR, sign=-1.0, non_nagetive=self.non_nagetive, snr_max=self.snr_max)

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        return self.fused_objf()(out, ref)


class WaTask(TimeDomainTask):
    """L1 / L2 distance between waveforms (summed over samples).  aps/task/sse.py:170-204 ("sse@wa")"""

    def __init__(self, nnet: nn.Module, objf: str = "L1", num_spks: int = 2, permute: bool = True,
                 weight: Optional[str] = None) -> None:
        super(WaTask, self).__init__(nnet, num_spks=num_spks, permute=permute, weight=weight,
                                     description="Using L1/L2 loss on waveform for training")
        self.l1 = objf == "L1"

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        d = out - ref
        return (d.abs() if self.l1 else d * d).sum(-1)

    def fused_objf(self):
        return self.objf                       # a plain callable: the PIT helpers evaluate it pair by pair


class FreqSaTask(Task):
    """Frequency-domain spectral approximation (to be inherited).  aps/task/sse.py:207-311.

    The polar STFTs of the mixture and of every reference come from the fused STFT kernel (`enh_transform.ctx(
    "forward_stft")`, one pass per signal, magnitude and phase written together); the reference magnitude (MSA /
    phase-sensitive / truncated), the masking product and the distance run as device tensor ops; permutations go
    through `hybrid_permu_objf` as in the reference.  The deep-clustering branch (`dpcl_weight > 0` with a network
    that has `dpcl_embed`) is not built."""

    def __init__(self, nnet: nn.Module, phase_sensitive: bool = False, truncated: float = -1, permute: bool = True,
                 masking: bool = True, num_spks: int = 2, description: str = "", dpcl_weight: float = 0,
                 weight: Optional[str] = None) -> None:
        super(FreqSaTask, self).__init__(nnet, ctx=nnet.enh_transform.ctx("forward_stft"), description=description)
        self.weight = list(map(float, weight.split(","))) if weight is not None else None
        if not masking and truncated > 0:
            raise ValueError("Conflict parameters: masksing = True while truncated > 0")
        if dpcl_weight > 0 and hasattr(nnet, "dpcl_embed") and num_spks > 1:
            raise NotImplementedError("FreqSaTask: the deep-clustering branch (dpcl_weight > 0) is not built")
        self.phase_sensitive = phase_sensitive
        self.truncated = truncated
        self.permute = permute
        self.masking = masking
        self.num_spks = num_spks

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        raise NotImplementedError

    def transform(self, tensor: th.Tensor) -> th.Tensor:
        raise NotImplementedError

    def _ref_mag(self, mix_in_polar: th.Tensor, ref_in_polar: th.Tensor) -> th.Tensor:
        """Reference magnitude: |S|, |S| max(cos(phase difference), 0) (PSA), min(., truncated |Y|) (tPSA)"""
        ref_mag, ref_pha = ref_in_polar[..., 0], ref_in_polar[..., 1]
        if self.phase_sensitive:
            ref_mag = ref_mag * th.clamp(th.cos(ref_pha - mix_in_polar[..., 1]), min=0)
        if self.truncated > 0:
            ref_mag = th.min(ref_mag, self.truncated * mix_in_polar[..., 0])
        return ref_mag

    def forward(self, egs: Dict) -> Dict:
        """egs: mix N x (C) x S, ref N x S or [N x S, ...]"""
        mix, ref = egs["mix"], egs["ref"]
        mask = self.nnet(mix)
        mix_in_polar = self.ctx(mix[:, 0] if mix.dim() == 3 else mix, return_polar=True)
        if isinstance(mask, th.Tensor):
            mask, ref = [mask], [ref]
        ref_mags = [self._ref_mag(mix_in_polar, self.ctx(r, return_polar=True)) for r in ref]
        out = [m * mix_in_polar[..., 0] for m in mask] if self.masking else mask
        loss = hybrid_permu_objf(out, ref_mags, self.objf, transform=self.transform, weight=self.weight,
                                 permute=self.permute, permu_num_spks=self.num_spks)
        return {"loss": loss.mean()}


class LinearFreqSaTask(FreqSaTask):
    """MSA / (t)PSA loss on the linear spectrogram, L1 or L2.  aps/task/sse.py:314-376 ("sse@freq_linear_sa")"""

    def __init__(self, nnet: nn.Module, phase_sensitive: bool = False, truncated: float = -1, permute: bool = True,
                 masking: bool = True, dpcl_weight: float = 0, num_spks: int = 2, objf: str = "L2",
                 weight: Optional[str] = None) -> None:
        super(LinearFreqSaTask, self).__init__(nnet, phase_sensitive=phase_sensitive, truncated=truncated,
                                               permute=permute, masking=masking, weight=weight, dpcl_weight=dpcl_weight,
                                               num_spks=num_spks,
                                               description="Using spectral approximation (MSA or tPSA) loss function")
        self.l1 = objf == "L1"

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        """out, ref: N x F x T -> N (mean over frames, sum over bins)"""
        d = out - ref
        return th.sum((d.abs() if self.l1 else d * d).mean(-1), -1)

    def transform(self, tensor: th.Tensor) -> th.Tensor:
        return tensor


class MelFreqSaTask(FreqSaTask):
    """L2 loss between (log-)mel spectrograms.  aps/task/sse.py:379-455 ("sse@freq_mel_sa")"""

    def __init__(self, nnet: nn.Module, phase_sensitive: bool = False, truncated: float = -1,
                 weight: Optional[str] = None, dpcl_weight: float = 0, permute: bool = True, num_spks: int = 2,
                 masking: bool = True, power_mag: bool = False, num_bins: int = 257, num_mels: int = 80,
                 mel_log: int = False, mel_scale: int = 1, mel_norm: bool = False, sr: int = 16000,
                 fmax: int = 8000) -> None:
        from ..transform.utils import mel_filter
        super(MelFreqSaTask, self).__init__(nnet, phase_sensitive=phase_sensitive, truncated=truncated, permute=permute,
                                            masking=masking, weight=weight, dpcl_weight=dpcl_weight, num_spks=num_spks,
                                            description="Using L2 loss of the mel features")
        mel = mel_filter(None, num_bins=num_bins, sr=sr, num_mels=num_mels, fmax=fmax, norm=mel_norm)
        self.mel = nn.Parameter(mel[..., None] * mel_scale, requires_grad=False)       # M x F x 1, as in the reference
        self.log = mel_log
        self.power_mag = power_mag

    def transform(self, tensor: th.Tensor) -> th.Tensor:
        """N x F x T -> N x M x T"""
        if self.power_mag:
            tensor = tensor**2
        mel = _mel_project(self.mel[..., 0].to(tensor.device), tensor)
        return th.log(1 + mel) if self.log else mel

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        d = out - ref
        return th.sum((d * d).mean(-1), -1)


class _MelLoss:
    """mel projection + L2 distance shared by MelFreqSaTask and MelTimeSaTask (sse.py:430-455, 660-681)"""

    def _init_mel(self, num_bins, num_mels, sr, fmax, mel_norm, mel_scale, mel_log, power_mag):
        from ..transform.utils import mel_filter
        mel = mel_filter(None, num_bins=num_bins, sr=sr, num_mels=num_mels, fmax=fmax, norm=mel_norm)
        self.mel = nn.Parameter(mel[..., None] * mel_scale, requires_grad=False)       # M x F x 1, as in the reference
        self.log = mel_log
        self.power_mag = power_mag


class TimeSaTask(Task):
    """Time-domain networks scored on STFT magnitudes (to be inherited).  aps/task/sse.py:458-541.

    The reference pre-emphasises `wav[:, 1:] -= a * wav[:, :-1]` IN PLACE, which also rewrites the caller's reference
    signals; here the pre-emphasised signal is a new tensor (same loss value, the inputs are left alone)."""

    def __init__(self, nnet: nn.Module, frame_len: int = 512, frame_hop: int = 256, center: bool = False,
                 window: str = "sqrthann", round_pow_of_two: bool = True, stft_normalized: bool = False,
                 pre_emphasis: float = 0, permute: bool = True, weight: Optional[str] = None, num_spks: int = 2,
                 description: str = "") -> None:
        from ..transform.utils import STFT
        sa_ctx = STFT(frame_len, frame_hop, window=window, center=center, round_pow_of_two=round_pow_of_two,
                      normalized=stft_normalized)
        super(TimeSaTask, self).__init__(nnet, ctx=sa_ctx, description=description)
        self.weight = list(map(float, weight.split(","))) if weight is not None else None
        self.permute = permute
        self.num_spks = num_spks
        self.pre_emphasis = pre_emphasis

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        raise NotImplementedError

    def transform(self, tensor: th.Tensor) -> th.Tensor:
        raise NotImplementedError

    def _stft_mag(self, wav: th.Tensor) -> th.Tensor:
        if self.pre_emphasis > 0:
            wav = th.cat([wav[:, :1], wav[:, 1:] - self.pre_emphasis * wav[:, :-1]], 1)
        return self.ctx(wav, return_polar=True)[..., 0]

    def forward(self, egs: Dict) -> Dict:
        """egs: mix N x (C) x S, ref N x S or [N x S, ...]"""
        mix, ref = egs["mix"], egs["ref"]
        spk = self.nnet(mix)
        if isinstance(spk, th.Tensor):
            spk, ref = [spk], [ref]
        spk_mag = [self._stft_mag(s) for s in spk]
        ref_mag = [self._stft_mag(r) for r in ref]
        loss = hybrid_permu_objf(spk_mag, ref_mag, self.objf, transform=self.transform, weight=self.weight,
                                 permute=self.permute, permu_num_spks=self.num_spks)
        return {"loss": th.mean(loss)}


class LinearTimeSaTask(TimeSaTask):
    """L1 / L2 distance between STFT magnitudes of the separated and reference waveforms.
    aps/task/sse.py:544-603 ("sse@time_linear_sa")"""

    def __init__(self, nnet: nn.Module, frame_len: int = 512, frame_hop: int = 256, center: bool = False,
                 window: str = "sqrthann", round_pow_of_two: bool = True, stft_normalized: bool = False,
                 permute: bool = True, weight: Optional[str] = None, num_spks: int = 2, objf: str = "L2") -> None:
        super(LinearTimeSaTask, self).__init__(nnet, frame_len=frame_len, frame_hop=frame_hop, window=window,
                                               center=center, round_pow_of_two=round_pow_of_two,
                                               stft_normalized=stft_normalized, permute=permute, num_spks=num_spks,
                                               weight=weight,
                                               description="Using L1/L2 loss on magnitude of the waveform")
        self.l1 = objf == "L1"

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        d = out - ref
        return th.sum((d.abs() if self.l1 else d * d).mean(-1), -1)

    def transform(self, tensor: th.Tensor) -> th.Tensor:
        return tensor


class MelTimeSaTask(TimeSaTask, _MelLoss):
    """L2 distance between (log-)mel spectrograms of the waveforms.  aps/task/sse.py:606-681 ("sse@time_mel_sa")"""

    def __init__(self, nnet: nn.Module, frame_len: int = 512, frame_hop: int = 256, window: str = "sqrthann",
                 center: bool = False, round_pow_of_two: bool = True, stft_normalized: bool = False,
                 permute: bool = True, weight: Optional[str] = None, num_spks: int = 2, num_bins: int = 257,
                 num_mels: int = 80, power_mag: bool = False, mel_log: bool = False, mel_scale: int = 1,
                 mel_norm: bool = False, sr: int = 16000, fmax: int = 7690) -> None:
        super(MelTimeSaTask, self).__init__(nnet, frame_len=frame_len, frame_hop=frame_hop, window=window,
                                            center=center, round_pow_of_two=round_pow_of_two,
                                            stft_normalized=stft_normalized, permute=permute, num_spks=num_spks,
                                            weight=weight,
                                            description="Using L2 loss on the mel features of the waveform")
        self._init_mel(num_bins, num_mels, sr, fmax, mel_norm, mel_scale, mel_log, power_mag)

    transform = MelFreqSaTask.transform
    objf = MelFreqSaTask.objf


class ComplexMappingTask(Task):
    """Complex spectral mapping: L1 / L2 on real and imaginary parts (+ magnitude).
    aps/task/sse.py:684-752 ("sse@complex_mapping")"""

    def __init__(self, nnet: nn.Module, num_spks: int = 2, weight: Optional[str] = None, permute: bool = True,
                 objf: str = "L1", add_magnitude_loss: bool = True) -> None:
        super(ComplexMappingTask, self).__init__(nnet, ctx=nnet.enh_transform.ctx("forward_stft"),
                                                 description="Using complex mapping function for training")
        self.weight = list(map(float, weight.split(","))) if weight is not None else None
        self.permute = permute
        self.num_spks = num_spks
        self.l1 = objf == "L1"
        self.add_magnitude_loss = add_magnitude_loss

    def _dist(self, a: th.Tensor, b: th.Tensor) -> th.Tensor:
        d = a - b
        return d.abs() if self.l1 else d * d

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        """out, ref: N x F x T x 2 -> N"""
        from .objf import EPSILON
        loss = self._dist(out[..., 0], ref[..., 0]) + self._dist(out[..., 1], ref[..., 1])
        if self.add_magnitude_loss:
            out_mag = th.sqrt(out[..., 0]**2 + out[..., 1]**2 + EPSILON)
            ref_mag = th.sqrt(ref[..., 0]**2 + ref[..., 1]**2 + EPSILON)
            loss = loss + self._dist(out_mag, ref_mag)
        return th.sum(loss.mean(-1), -1)

    def forward(self, egs: Dict) -> Dict:
        mix, ref = egs["mix"], egs["ref"]
        out = self.nnet(mix)
        if isinstance(out, th.Tensor):
            out, ref = [out], [ref]
        ref = [self.ctx(r, return_polar=False) for r in ref]
        loss = hybrid_permu_objf(out, ref, self.objf, weight=self.weight, permute=self.permute,
                                 permu_num_spks=self.num_spks)
        return {"loss": th.mean(loss)}


class ComplexMaskingTask(ComplexMappingTask):
    """Complex ratio masks: distance after masking, or to the compressed ideal mask.
    aps/task/sse.py:755-841 ("sse@complex_masking")"""

    def __init__(self, nnet: nn.Module, num_spks: int = 2, weight: Optional[str] = None, permute: bool = True,
                 compress_param=(10, 0.1, -100), compress_masks: bool = False, objf: str = "L2") -> None:
        super(ComplexMaskingTask, self).__init__(nnet, num_spks=num_spks, weight=weight, permute=permute, objf=objf,
                                                 add_magnitude_loss=False)
        self.k, self.c, self.lower_bound = compress_param
        self.compress_masks = compress_masks

    def _compress_mask(self, mix_stft: th.Tensor, ref: th.Tensor) -> th.Tensor:
        """compressed complex ratio mask in [-k, k] (sse.py:783-796)"""
        from .objf import EPSILON
        ref_stft = self.ctx(ref, return_polar=False)
        denominator = th.sum(mix_stft**2, -1) + EPSILON
        real = mix_stft[..., 0] * ref_stft[..., 0] + mix_stft[..., 1] * ref_stft[..., 1]
        imag = mix_stft[..., 0] * ref_stft[..., 1] - mix_stft[..., 1] * ref_stft[..., 0]
        # as written in the reference: [N, F, T, 2] / [N, F, T] only broadcasts for special shapes and raises otherwise
        crm = th.stack([real, imag], -1) / denominator
        exp = th.exp(-self.c * th.clamp_min(crm, self.lower_bound))
        return self.k * (1 - exp) / (1 + exp)

    @staticmethod
    def _complex_tf_mask(mix_stft: th.Tensor, mask: th.Tensor) -> th.Tensor:
        real = mix_stft[..., 0] * mask[..., 0] - mix_stft[..., 1] * mask[..., 1]
        imag = mix_stft[..., 0] * mask[..., 1] + mix_stft[..., 1] * mask[..., 0]
        return th.stack([real, imag], -1)

    def forward(self, egs: Dict) -> Dict:
        ref = egs["ref"]
        out = self.nnet(egs["mix"])
        if isinstance(out, th.Tensor):
            out, ref = [out], [ref]
        mix = self.ctx(egs["mix"], return_polar=False)
        if self.compress_masks:
            ref = [self._compress_mask(mix, r) for r in ref]
        else:
            ref = [self.ctx(r, return_polar=False) for r in ref]
            out = [self._complex_tf_mask(mix, o) for o in out]
        loss = hybrid_permu_objf(out, ref, self.objf, weight=self.weight, permute=self.permute,
                                 permu_num_spks=self.num_spks)
        return {"loss": th.mean(loss)}
