"""`EnhTransform` (class name `FeatureTransform`, registry key "enh") with the reference's surface
(/root/reference/aps/transform/enh.py:387-613): `encode` (multi-channel STFT, kernel F2), `decode`
(iSTFT, kernel F3), `forward` (reference-channel magnitude chain + IPD from a packed STFT, kernels
F1b + IPD), `ctx`, `num_frames`, `dim`, and the attributes `.forward_stft`, `.inverse_stft`,
`.mag_transform`, `.ipd_transform`, `.feats_dim`, `.frame_len`, `.frame_hop`.

Not provided (geometry-specific, not on the named path): `DfTransform`, `FixedBeamformer`.
"""
from typing import List, Optional, Tuple

import torch as th
import torch.nn as nn

from .. import _lib
from .asr import (AsrReturnType, FeatureTransform as AsrTransform, TFTransposeTransform, _Layer, _match_tail,
                  check_valid, fused_spec_features)
from .utils import EPSILON, STFT, iSTFT


class RefChannelTransform(_Layer):
    """Pick the reference channel of an N x C x ... tensor (enh.py:21-49)."""

    def __init__(self, ref_channel: int = 0, input_dim: int = 4) -> None:
        super().__init__()
        self.ref_channel, self.input_dim = ref_channel, input_dim

    def extra_repr(self) -> str:
        return f"ref_channel={self.ref_channel}"

    def forward(self, inp: th.Tensor) -> th.Tensor:
        if inp.dim() != self.input_dim or self.ref_channel < 0:
            return inp
        return inp[:, self.ref_channel]


class PhaseTransform(_Layer):
    """[real, imag] -> angle (enh.py:52-76)."""

    def __init__(self, dim: int = -1):
        super().__init__()
        self.dim = dim

    def extra_repr(self) -> str:
        return f"dim={self.dim}"

    def forward(self, inp: th.Tensor) -> th.Tensor:
        return th.atan2(th.select(inp, self.dim, 1), th.select(inp, self.dim, 0))


class IpdTransform(_Layer):
    """Inter-channel phase differences for the pairs "l,r;l,r;..." (enh.py:79-143)."""

    def __init__(self, ipd_index: str = "1,0", cos: bool = True, sin: bool = False) -> None:
        super().__init__()
        pairs = [tuple(map(int, p.split(","))) for p in ipd_index.split(";")]
        self.index_l = [p[0] for p in pairs]
        self.index_r = [p[1] for p in pairs]
        self.ipd_index = ipd_index
        self.cos, self.sin = cos, sin
        self.num_pairs = len(pairs) * 2 if cos and sin else len(pairs)
        self._idx = None

    def extra_repr(self) -> str:
        return f"ipd_index={self.ipd_index}, cos={self.cos}, sin={self.sin}"

    def device_index(self, dev: th.device):
        if self._idx is None or self._idx[0].device != dev:
            self._idx = (th.tensor(self.index_l, dtype=th.int32, device=dev),
                         th.tensor(self.index_r, dtype=th.int32, device=dev))
        return self._idx

    def forward(self, p: th.Tensor) -> th.Tensor:
        """p: phase N x C x T x F -> N x T x MF (layer-by-layer form; EnhTransform.forward uses the kernel)."""
        if p.dim() not in (3, 4):
            raise RuntimeError(f"{self.__class__.__name__} expect 3/4D tensor, but got {p.dim():d} instead")
        if p.dim() == 3:
            p = p.unsqueeze(0)
        N, C, T, _ = p.shape
        assert C != 1
        p = p.transpose(1, 2)
        dif = p[..., self.index_l, :] - p[..., self.index_r, :]
        if not self.cos:
            raise RuntimeError("IpdTransform(cos=False) is not usable in the reference either (enh.py:137-139)")
        ipd = th.cos(dif)
        if self.sin:
            ipd = th.cat([ipd, th.sin(dif)], 2)
        return ipd.reshape(N, T, -1)


def ipd_features(packed: th.Tensor, ipd: IpdTransform, out: Optional[th.Tensor] = None, col0: int = 0) -> th.Tensor:
    """IPD kernel launch: packed N x C x F x T x 2 -> columns [col0, col0 + num_pairs*F) of `out`."""
    dev = _lib.require_cuda(packed, "the packed STFT")
    if packed.dim() != 5:
        raise RuntimeError(f"IPD features need a multi-channel packed STFT N x C x F x T x 2, got {packed.dim()}D")
    if not ipd.cos:
        raise RuntimeError("IpdTransform(cos=False) is not usable in the reference either (enh.py:137-139)")
    x = packed.detach().float().contiguous()
    N, C, F, T, _ = x.shape
    assert C != 1
    if max(ipd.index_l + ipd.index_r) >= C or min(ipd.index_l + ipd.index_r) < 0:
        raise IndexError(f"ipd_index {ipd.ipd_index} out of range for {C} channels")
    P = len(ipd.index_l)
    cols = P * F * (2 if ipd.sin else 1)
    if out is None:
        out = th.empty((N, T, cols), dtype=th.float32, device=dev)
        col0 = 0
    il, ir = ipd.device_index(dev)
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_ipd_fwd(x.data_ptr(), N, C, F, T, il.data_ptr(), ir.data_ptr(), P,
                                                int(ipd.sin), out.data_ptr(), out.shape[-1], col0,
                                                _lib.stream_ptr(dev)))
    return out


class FeatureTransform(nn.Module):
    """Feature transform for enhancement / separation — arguments as in the reference (enh.py:388-468)."""

    def __init__(self,
                 feats: str = "spectrogram-log-cmvn",
                 frame_len: int = 512,
                 frame_hop: int = 256,
                 window: str = "sqrthann",
                 round_pow_of_two: bool = True,
                 stft_normalized: bool = False,
                 stft_mode: str = "librosa",
                 center: bool = False,
                 ref_channel: int = 0,
                 use_power: bool = False,
                 sr: int = 16000,
                 log_lower_bound: float = 0,
                 num_mels: int = 80,
                 mel_matrix: str = "",
                 mel_coeff_norm: bool = False,
                 min_freq: int = 0,
                 max_freq: Optional[int] = None,
                 num_ceps: int = 13,
                 lifter: float = 0,
                 aug_prob: float = 0,
                 aug_adaptive_args: Tuple[int] = (0, 0),
                 aug_mask_zero: bool = True,
                 aug_time_args: Tuple[int] = (40, 1),
                 aug_freq_args: Tuple[int] = (30, 1),
                 norm_mean: bool = True,
                 norm_var: bool = True,
                 norm_per_band: bool = True,
                 gcmvn: str = "",
                 subsampling_factor: int = 1,
                 lctx: int = 1,
                 rctx: int = 1,
                 delta_ctx: int = 2,
                 delta_order: int = 2,
                 delta_as_channel: bool = False,
                 requires_grad: bool = False,
                 ipd_index: str = "",
                 cos_ipd: bool = True,
                 sin_ipd: bool = False,
                 eps: float = EPSILON) -> None:
        super().__init__()
        self.frame_len, self.frame_hop = frame_len, frame_hop
        self.stft_kwargs = dict(mode=stft_mode, window=window, center=center, normalized=stft_normalized,
                                round_pow_of_two=round_pow_of_two)
        self.forward_stft = self.ctx(name="forward_stft")
        self.inverse_stft = self.ctx(name="inverse_stft")
        tokens = feats.split("-")
        dim = 0
        mag_tokens = "-".join(t for t in tokens if t != "ipd")
        if mag_tokens:
            # NOTE: like the reference, pre_emphasis / audio_norm / eps stay at AsrTransform's defaults here
            asr = AsrTransform(feats=mag_tokens, frame_len=frame_len, frame_hop=frame_hop, window=window,
                               round_pow_of_two=round_pow_of_two, stft_normalized=stft_normalized,
                               stft_mode=stft_mode, center=center, use_power=use_power, sr=sr,
                               log_lower_bound=log_lower_bound, num_mels=num_mels, mel_matrix=mel_matrix,
                               mel_coeff_norm=mel_coeff_norm, min_freq=min_freq, max_freq=max_freq,
                               num_ceps=num_ceps, lifter=lifter, aug_prob=aug_prob,
                               aug_adaptive_args=aug_adaptive_args, aug_mask_zero=aug_mask_zero,
                               aug_time_args=aug_time_args, aug_freq_args=aug_freq_args, norm_mean=norm_mean,
                               norm_var=norm_var, norm_per_band=norm_per_band, gcmvn=gcmvn,
                               subsampling_factor=subsampling_factor, lctx=lctx, rctx=rctx, delta_ctx=delta_ctx,
                               delta_order=delta_order, delta_as_channel=delta_as_channel,
                               requires_grad=requires_grad)
            if asr.spectra_index == -1:
                raise RuntimeError("Now only support spectrogram/mfcc/fbank features")
            dim = asr.dim()
            # drop the SpectrogramTransform: this chain starts from the packed STFT (enh.py:525-529)
            self.mag_transform = nn.Sequential(RefChannelTransform(ref_channel=ref_channel, input_dim=5),
                                               *list(asr.transform[1:]))
        else:
            self.mag_transform = None
        if any(t == "ipd" for t in tokens) and ipd_index:
            self.ipd_transform = nn.Sequential(PhaseTransform(dim=-1), TFTransposeTransform(),
                                               IpdTransform(ipd_index=ipd_index, cos=cos_ipd, sin=sin_ipd))
            pairs = len(ipd_index.split(";"))
            dim += pairs * (2 if cos_ipd and sin_ipd else 1) * self.forward_stft.num_bins
        else:
            self.ipd_transform = None
        self.feats_dim = dim

    def dim(self) -> int:
        return self.feats_dim

    def ctx(self, name: str = "forward_stft") -> nn.Module:
        """A fresh STFT / iSTFT layer with this transform's framing (enh.py:553-560)."""
        table = {"forward_stft": STFT, "inverse_stft": iSTFT}
        if name not in table:
            raise ValueError(f"Unknown task context: {name}")
        return table[name](self.frame_len, self.frame_hop, **self.stft_kwargs)

    def num_frames(self, wav_len: Optional[th.Tensor]) -> Optional[th.Tensor]:
        return None if wav_len is None else self.forward_stft.num_frames(wav_len)

    def encode(self, wav_pad: th.Tensor, wav_len: Optional[th.Tensor]) -> AsrReturnType:
        """N x (C) x S -> (packed N x (C) x F x T x 2, num_frames) (enh.py:571-584)."""
        return self.forward_stft(wav_pad, return_polar=False), self.num_frames(wav_len)

    def decode(self, packed: List[th.Tensor]) -> List[th.Tensor]:
        """[N x F x T x 2, ...] -> [N x S, ...] (enh.py:586-593)."""
        return [self.inverse_stft(p, return_polar=False) for p in packed]

    def forward(self, packed: th.Tensor) -> th.Tensor:
        """packed N x (C) x F x T x 2 -> spectral (+ spatial) features N x T x D (enh.py:595-613)."""
        _lib.require_cuda(packed, "EnhTransform input")
        ipd = self.ipd_transform[2] if self.ipd_transform is not None else None
        ipd_cols = 0
        if ipd is not None:
            ipd_cols = len(ipd.index_l) * self.forward_stft.num_bins * (2 if ipd.sin else 1)
        out = None
        if self.mag_transform is not None:
            layers = list(self.mag_transform)
            tail = _match_tail(layers, 1)
            ref = layers[0]
            if tail is None or not isinstance(ref, RefChannelTransform):
                raise RuntimeError("aps_b200: unsupported magnitude chain for EnhTransform")
            ref_ch = ref.ref_channel if (packed.dim() == ref.input_dim and ref.ref_channel >= 0) else None
            if packed.dim() == 5 and ref_ch is None:
                raise RuntimeError("aps_b200: ref_channel < 0 (all channels) is not implemented for EnhTransform")
            rest = layers[tail[5]:]
            fuse_ipd = ipd is not None and not rest and not (tail[3] is not None and tail[3].gmean is None
                                                              and not tail[3].per_band)
            out, D = fused_spec_features(packed, ref_ch or 0, tail, extra_cols=ipd_cols if fuse_ipd else 0)
            for lay in rest:
                out = lay(out)
            if ipd is not None:
                if fuse_ipd:
                    ipd_features(packed, ipd, out=out, col0=D)
                else:
                    out = th.cat([out, ipd_features(packed, ipd)], -1)
        elif ipd is not None:
            out = ipd_features(packed, ipd)
        if out is None:
            raise RuntimeError("EnhTransform has neither spectral nor spatial features configured")
        return check_valid(out, None)[0]
