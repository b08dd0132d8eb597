"""Frequency-domain Transformer separation / enhancement model (`sse@freq_xfmr`) with the constructor, methods and
`state_dict` layout of /root/reference/aps/sse/bss/transformer.py:16-128: STFT features -> TransformerEncoder (linear
projection front, `output_proj = num_bins * num_spks`) -> mask non-linearity -> per-speaker masks `N x F x T`.

The encoder is `aps_b200.asr.transformer.TransformerEncoder` (the tcgen05 GEMM engine, the register-tiled attention,
fused LayerNorm kernels): every parameter sits under `xfmr.*` exactly as in the reference, so checkpoints load with
strict=True.  `mask` is kept as an (empty-state) `nn.Sequential` for the same reason.  Inference only.
"""
from typing import Dict, List, Optional

import torch as th
import torch.nn as nn

from ...asr.transformer import TransformerEncoder

_NON_LINEAR = {"relu": th.relu, "sigmoid": th.sigmoid, "softmax": None}   # aps/sse/base.py:112-156 ("common" set)


class FreqXfmr(nn.Module):
    """Arguments as in aps/sse/bss/transformer.py:22-36."""

    def __init__(self, enh_transform: Optional[nn.Module] = None, input_size: int = 257, num_spks: int = 2,
                 num_bins: int = 257, rctx: int = -1, lctx: int = -1, arch: str = "xfmr", pose: str = "rel",
                 arch_kwargs: Dict = {}, pose_kwargs: Dict = {}, proj_kwargs: Dict = {}, num_layers: int = 6,
                 non_linear: str = "sigmoid", training_mode: str = "freq") -> None:
        super().__init__()
        assert enh_transform is not None
        assert training_mode in ("freq", "time")
        if non_linear not in _NON_LINEAR:
            raise ValueError(f"Unsupported nonlinear: {non_linear}")
        self.enh_transform = enh_transform
        self.training_mode = training_mode
        self.xfmr = TransformerEncoder(arch, input_size, output_proj=num_bins * num_spks, num_layers=num_layers,
                                       chunk_size=1, lctx=lctx, rctx=rctx, proj="linear", proj_kwargs=proj_kwargs,
                                       pose=pose, pose_kwargs=pose_kwargs, arch_kwargs=arch_kwargs)
        self.mask = nn.Sequential()            # MaskNonLinear + transpose in the reference: no parameters
        self.non_linear = non_linear
        self.num_spks, self.num_bins = num_spks, num_bins

    def _tf_mask(self, feats: th.Tensor, num_spks: int) -> List[th.Tensor]:
        """[N x F x T, ...] — transformer.py:57-67 (softmax is taken over the speaker axis, sse/base.py:140-150)."""
        out, _ = self.xfmr(feats, None)                                  # N x T x (S*F)
        if self.non_linear == "softmax":
            N, T, _ = out.shape
            out = th.softmax(out.view(N, T, num_spks, -1), 2).view(N, T, -1)
        else:
            out = _NON_LINEAR[self.non_linear](out)
        return list(th.chunk(out.transpose(1, 2), num_spks, 1))

    def mask_predict(self, feats: th.Tensor) -> th.Tensor:
        """feats N x T x F -> masks N x F x T (or S x N x F x T) — transformer.py:117-128."""
        masks = th.stack(self._tf_mask(feats, self.num_spks))
        return masks[0] if self.num_spks == 1 else masks

    def _infer(self, mix: th.Tensor, mode: str = "freq"):
        """transformer.py:69-84"""
        stft, _ = self.enh_transform.encode(mix, None)
        masks = self._tf_mask(self.enh_transform(stft), self.num_spks)
        if mode == "time":
            ref = stft[:, 0] if stft.dim() == 5 else stft
            packed = self.enh_transform.decode([ref * m.unsqueeze(-1) for m in masks])     # aps/sse/base.py:23-49
        else:
            packed = masks
        return packed[0] if self.num_spks == 1 else packed

    def infer(self, mix: th.Tensor, mode: str = "time"):
        if mix.dim() != 1:
            raise RuntimeError(f"FreqXfmr expects 1D tensor (inference), got {mix.dim()} instead")
        with th.no_grad():
            sep = self._infer(mix[None, :], mode=mode)
            return sep[0] if self.num_spks == 1 else [s[0] for s in sep]

    def forward(self, s: th.Tensor):
        if s.dim() != 2:
            raise RuntimeError(f"FreqXfmr expects 2D tensor (training), got {s.dim()} instead")
        return self._infer(s, mode=self.training_mode)
