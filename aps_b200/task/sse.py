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
