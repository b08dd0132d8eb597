"""Time-domain separation tasks (network forward + Si-SNR / SNR value) on the fused objective kernel.

Same constructor arguments, `forward(egs) -> {"loss": scalar}` contract and registry aliases
("sse@sisnr", "sse@snr") as aps/task/sse.py:60-167.  Forward (evaluation) only: the loss tensor
carries no autograd graph, training through it is out of scope.
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
        mel = th.matmul(self.mel[..., 0].to(tensor.device), tensor)
        return th.log(1 + mel) if self.log else mel

    def objf(self, out: th.Tensor, ref: th.Tensor) -> th.Tensor:
        d = out - ref
        return th.sum((d * d).mean(-1), -1)
