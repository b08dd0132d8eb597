"""`AsrTransform` (class name `FeatureTransform`, registry key "asr") with the reference's surface
(/root/reference/aps/transform/asr.py:784-1033) on top of the fused sm_100a feature kernel.

The token string still builds an indexable `nn.Sequential` of small layer modules — same order,
same parameter names/shapes, so `state_dict()` keys (`transform.<i>.K|w|filters|...`) and the
attributes other code reads (`.transform`, `.spectra_index`, `.perturb_index`, `.feats_dim`,
`.subsampling_factor`) are unchanged — but `forward` does not run them one by one: `run_chain`
recognises the run

    [Rescale] [PreEmphasis] Spectrogram Magnitude TFTranspose Power [Mel] [Log] [Cmvn]

and issues ONE kernel for it (csrc/frontend.cu, F1).  The same planner serves `EnhTransform`, whose
magnitude chain starts from a packed STFT (csrc/specfeat.cu, F1b).  Layers outside that run
(speed perturbation, DCT, SpecAug, splice, delta — SURVEY.md §8f "next" rows) execute as plain
device tensor ops for now.
"""
import math
import random
import warnings
from typing import List, Optional, Tuple, Union

import numpy as np
import torch as th
import torch.nn as nn
import torch.nn.functional as tf

from .. import _lib
from .utils import EPSILON, MAX_INT16, STFT, mel_filter, stft_forward

AsrReturnType = Union[th.Tensor, Optional[th.Tensor]]


def check_valid(feature: th.Tensor, num_frames: Optional[th.Tensor],
                nan_count: Optional[th.Tensor] = None) -> Tuple[th.Tensor, Optional[th.Tensor]]:
    """NaN guard + trim to the longest utterance (asr.py:33-53).  Raises ValueError on NaNs and
    RuntimeError when the feature matrix is shorter than `num_frames` claims.  `nan_count` is the
    counter the fused kernel filled while writing `feature` (saves re-reading the features)."""
    shape = feature.shape
    num_nans = int(nan_count) if nan_count is not None else int(th.isnan(feature).sum())
    if num_nans:
        raise ValueError(f"Detect {num_nans} NANs in feature matrices, shape = {shape}...")
    if num_frames is not None:
        longest = int(num_frames.max())
        if shape[-2] < longest:
            raise RuntimeError(f"feats shape: {shape[-2]} x {shape[-1]}, num_frames = {num_frames.tolist()}")
        if shape[-2] > longest:
            feature = feature[..., :longest, :]
    return feature, num_frames


class _Layer(nn.Module):
    """Small shared base: every layer says whether TorchScript export keeps it (asr.py `exportable`)."""
    _exportable = True

    def exportable(self) -> bool:
        return self._exportable


# ------------------------------------------------------------------------------------ waveform layers
class RescaleTransform(_Layer):
    """x -> round(x * rescale) (asr.py:56-84); folded into the staging step of F1/F2."""
    _exportable = False

    def __init__(self, rescale: float = MAX_INT16 * 1.0) -> None:
        super().__init__()
        self.rescale = rescale

    def extra_repr(self) -> str:
        return f"rescale={self.rescale}"

    def forward(self, wav: th.Tensor) -> th.Tensor:
        return th.round(wav * self.rescale)


class PreEmphasisTransform(_Layer):
    """Utterance-level pre-emphasis, IN PLACE on the caller's tensor like the reference (asr.py:87-113, Q9)."""
    _exportable = False

    def __init__(self, pre_emphasis: float = 0) -> None:
        super().__init__()
        self.pre_emphasis = pre_emphasis

    def extra_repr(self) -> str:
        return f"pre_emphasis={self.pre_emphasis}"

    def forward(self, wav: th.Tensor) -> th.Tensor:
        if self.pre_emphasis > 0:
            wav[..., 1:] = wav[..., 1:] - self.pre_emphasis * wav[..., :-1]
        return wav


def speed_perturb_filter(src_sr: int, dst_sr: int, cutoff_ratio: float = 0.95, num_zeros: int = 64) -> th.Tensor:
    """Polyphase windowed-sinc resampling bank [dst, src, K] (utils.py:159-190)."""
    if src_sr == dst_sr:
        raise ValueError(f"src_sr should not be equal to dst_sr: {src_sr}/{dst_sr}")
    g = math.gcd(src_sr, dst_sr)
    src, dst = src_sr // g, dst_sr // g
    if src == 1 or dst == 1:
        raise ValueError("do not support integer downsample/upsample")
    zpb = min(src, dst) * cutoff_ratio
    half = 1 + int(num_zeros / zpb)
    t = (np.arange(dst)[:, None, None] / float(dst) - np.arange(src)[None, :, None] / float(src) -
         np.arange(2 * half + 1)[None, None, :] + half)
    taper = np.where(np.abs(t / half) < 1, 0.5 + 0.5 * np.cos(t / half * math.pi), 0.0)
    return th.tensor(np.sinc(t * zpb) * taper * zpb / float(src), dtype=th.float32)


class SpeedPerturbTransform(_Layer):
    """Train-time per-utterance speed perturbation (asr.py:116-195; "next" row f2: device tensor ops)."""
    _exportable = False

    def __init__(self, sr: int = 16000, perturb: str = "0.9,1.0,1.1") -> None:
        super().__init__()
        self.sr = sr
        self.factor_str = perturb
        rates = [int(f * sr) for f in map(float, perturb.split(","))]
        if not rates:
            raise ValueError("No perturb options for doing speed perturb")
        if sr not in rates:
            raise ValueError(f"We should keep 1.0 in perturb options: {perturb}")
        self.weights = nn.ParameterList(
            [nn.Parameter(speed_perturb_filter(sr, fs), requires_grad=False) for fs in rates if fs != sr])
        shp = [w.shape for w in self.weights]
        self.register_buffer("src_sr", th.tensor([s[1] for s in shp] + [1], dtype=th.int64))
        self.register_buffer("dst_sr", th.tensor([s[0] for s in shp] + [1], dtype=th.int64))
        self.last_choice = None

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(sr={self.sr}, factor={self.factor_str})"

    def output_length(self, inp_len: Optional[th.Tensor]) -> Optional[th.Tensor]:
        if self.last_choice is None or inp_len is None:
            return inp_len
        c = self.last_choice
        src, dst = self.src_sr.to(inp_len.device)[c.to(inp_len.device)], self.dst_sr.to(inp_len.device)[c.to(inp_len.device)]
        return th.div(inp_len, src, rounding_mode="trunc") * dst

    @staticmethod
    def _resample(wav: th.Tensor, weight: th.Tensor) -> th.Tensor:
        # augment.py:85-109 (perturb_speed): block the signal by src_sr, one conv1d over blocks
        _, src, K = weight.shape
        n, S = wav.shape
        blocks = S // src
        if blocks == 0:
            raise RuntimeError(f"Input wav is too short to be perturbed, length = {S}")
        x = wav[:, :blocks * src].view(n, blocks, src).transpose(1, 2)
        y = tf.conv1d(x, weight, padding=(K - 1) // 2)
        return y.transpose(1, 2).contiguous().view(n, -1)

    def forward(self, wav: th.Tensor) -> th.Tensor:
        self.last_choice = None
        if not self.training:
            return wav
        if wav.dim() != 2:
            raise RuntimeError(f"Now only supports 2D tensor, got {wav.dim()}")
        choice = th.randint(0, len(self.weights) + 1, (wav.shape[0],))
        self.last_choice = choice
        if wav.is_cuda:
            # one launch for the whole batch: every utterance runs the polyphase filter of its own choice
            # (csrc/featops.cu speed_perturb_kernel) and the result comes back zero padded to the longest one
            from .. import ops
            return ops.speed_perturb(wav, choice, list(self.weights))
        outs = []
        for i, c in enumerate(choice.tolist()):
            outs.append(wav[i] if c == len(self.weights) else self._resample(wav[i:i + 1], self.weights[c])[0])
        pad = th.zeros([wav.shape[0], max(o.shape[-1] for o in outs)], device=wav.device)
        for i, o in enumerate(outs):
            pad[i, :o.shape[-1]] = o
        return pad


# ------------------------------------------------------------------------------------ spectral layers
class TFTransposeTransform(_Layer):
    """Swap the last two axes (asr.py:196-223)."""

    def __init__(self, axis1: int = -1, axis2: int = -2) -> None:
        super().__init__()
        self.axis1, self.axis2 = axis1, axis2

    def extra_repr(self) -> str:
        return f"axis1={self.axis1}, axis2={self.axis2}"

    def forward(self, tensor: th.Tensor) -> th.Tensor:
        return tensor.transpose(-1, -2)


class SpectrogramTransform(STFT):
    """STFT layer of the feature chain: N x (C) x S -> N x (C) x F x T x 2 (asr.py:225-277)."""

    def __init__(self, frame_len: int, frame_hop: int, center: bool = False, window: str = "hamm",
                 round_pow_of_two: bool = True, normalized: bool = False, pre_emphasis: float = 0.97,
                 onesided: bool = True, mode: str = "librosa") -> None:
        super().__init__(frame_len, frame_hop, center=center, window=window, round_pow_of_two=round_pow_of_two,
                         pre_emphasis=pre_emphasis, normalized=normalized, onesided=onesided, mode=mode)

    def dim(self) -> int:
        return self.num_bins

    def exportable(self) -> bool:
        return False

    def forward(self, wav: th.Tensor) -> th.Tensor:
        return super().forward(wav, return_polar=False)


class MagnitudeTransform(_Layer):
    """sqrt(sum_dim x^2 + eps) (asr.py:280-303)."""

    def __init__(self, dim: int = -1, eps: float = 0):
        super().__init__()
        self.dim, self.eps = dim, eps

    def extra_repr(self) -> str:
        return f"dim={self.dim}, eps={self.eps}"

    def forward(self, inp: th.Tensor) -> th.Tensor:
        return th.sqrt(th.sum(inp**2, self.dim) + self.eps)


class AbsTransform(_Layer):
    """|x| for real tensors or ComplexTensor-like objects (asr.py:306-332)."""

    def __init__(self, eps: float = 1e-6) -> None:
        super().__init__()
        self.eps = eps

    def extra_repr(self) -> str:
        return f"eps={self.eps:.3e}"

    def forward(self, tensor):
        if not isinstance(tensor, th.Tensor):
            tensor = tensor + self.eps
        return tensor.abs()


class PowerTransform(_Layer):
    """x ** power (asr.py:335-357)."""

    def __init__(self, power: float = 2) -> None:
        super().__init__()
        self.power = power

    def extra_repr(self) -> str:
        return f"power={self.power}"

    def forward(self, tensor: th.Tensor) -> th.Tensor:
        return tensor**self.power


class MelTransform(_Layer):
    """Mel filterbank layer; `filters` [num_mels, num_bins] stays a Parameter (asr.py:360-428)."""

    def __init__(self, frame_len: int, round_pow_of_two: bool = True, sr: int = 16000, num_mels: int = 80,
                 fmin: float = 0.0, fmax: Optional[float] = None, mel_matrix: str = "", coeff_norm: bool = False,
                 requires_grad: bool = False) -> None:
        super().__init__()
        if mel_matrix:
            filters = th.load(mel_matrix)
        else:
            filters = mel_filter(frame_len, round_pow_of_two=round_pow_of_two, sr=sr, num_mels=num_mels,
                                 fmax=fmax, fmin=fmin, norm=coeff_norm)
        self.num_mels, self.num_bins = filters.shape
        self.filters = nn.Parameter(filters, requires_grad=requires_grad)
        self.fmin = fmin
        self.fmax = sr // 2 if fmax is None else fmax
        self.init = mel_matrix if mel_matrix else "librosa"
        self._bands = None

    def dim(self) -> int:
        return self.num_mels

    def extra_repr(self) -> str:
        shape = self.filters.shape
        return f"fmin={self.fmin}, fmax={self.fmax}, mel_filter={shape[0]}x{shape[1]}, init={self.init}"

    def bands(self, dev: th.device):
        """Banded form (start, len, weights[M, stride]) of `filters` for the kernels, rebuilt whenever
        the parameter changes (load_state_dict, optimiser step).  One small D2H copy per rebuild."""
        f = self.filters
        key = (f.data_ptr(), f._version, str(dev))
        if self._bands is None or self._bands[0] != key:
            dense = f.detach().float().cpu().numpy()
            M, F = dense.shape
            nz = dense != 0
            start = np.zeros(M, dtype=np.int32)
            length = np.zeros(M, dtype=np.int32)
            for m in range(M):
                idx = np.nonzero(nz[m])[0]
                if idx.size:
                    start[m], length[m] = idx[0], idx[-1] - idx[0] + 1
            stride = int(max(1, length.max())) | 1            # odd: conflict-free shared-memory rows
            w = np.zeros((M, stride), dtype=np.float32)
            for m in range(M):
                w[m, :length[m]] = dense[m, start[m]:start[m] + length[m]]
            pack = (th.from_numpy(start).to(dev), th.from_numpy(length).to(dev), th.from_numpy(w).to(dev), stride)
            self._bands = (key, pack)
        return self._bands[1]

    def forward(self, linear: th.Tensor) -> th.Tensor:
        if linear.dim() not in (3, 4):
            raise RuntimeError(f"MelTransform expect 3/4D tensor, but got {linear.dim()} instead")
        if linear.is_cuda:
            from .. import ops
            return ops.project_rows(linear, self.filters)       # unfused chains: this package's GEMM, not a library one
        return tf.linear(linear, self.filters, bias=None)


class LogTransform(_Layer):
    """log(clamp(x, eps)) or log(lower_bound + x) (asr.py:431-464)."""

    def __init__(self, eps: float = 1e-5, lower_bound: float = 0.0) -> None:
        super().__init__()
        self.eps, self.lower_bound = eps, lower_bound

    def dim_scale(self) -> int:
        return 1

    def extra_repr(self) -> str:
        return f"eps={self.eps:.3e}, lower_bound={self.lower_bound}"

    def forward(self, linear: th.Tensor) -> th.Tensor:
        x = self.lower_bound + linear if self.lower_bound > 0 else th.clamp(linear, min=self.eps)
        return th.log(x)


class DiscreteCosineTransform(_Layer):
    """Orthonormal DCT-II (+ cepstral lifter) for MFCC (asr.py:467-517; "next" row f3)."""

    def __init__(self, num_ceps: int = 13, num_mels: int = 40, lifter: float = 0) -> None:
        super().__init__()
        self.lifter, self.num_ceps = lifter, num_ceps
        n = np.arange(num_mels, dtype=np.float64)
        k = np.arange(num_ceps, dtype=np.float64)[:, None]
        mat = np.cos(math.pi * (2 * n + 1) * k / (2 * num_mels)) * math.sqrt(2.0 / num_mels)
        mat[0] *= math.sqrt(0.5)
        self.dct = nn.Parameter(th.from_numpy(mat.astype(np.float32)), requires_grad=False)
        if lifter > 0:
            lift = 1 + lifter * 0.5 * th.sin(math.pi * th.arange(1, 1 + num_ceps) / lifter)
            self.cepstral_lifter = nn.Parameter(lift, requires_grad=False)
        else:
            self.cepstral_lifter = None

    def dim(self) -> int:
        return self.num_ceps

    def extra_repr(self) -> str:
        return "cepstral_lifter={0}, dct={1[0]}x{1[1]}".format(self.lifter, self.dct.shape)

    def forward(self, log_mel: th.Tensor) -> th.Tensor:
        if log_mel.is_cuda:
            from .. import ops
            out = ops.project_rows(log_mel, self.dct)
        else:
            out = tf.linear(log_mel, self.dct, bias=None)
        return out if self.cepstral_lifter is None else out * self.cepstral_lifter


class CmvnTransform(_Layer):
    """Utterance (per-frame "per_band" / "all band") or global mean-variance normalisation (asr.py:520-618)."""

    def __init__(self, norm_mean: bool = True, norm_var: bool = True, per_band: bool = True, dim: int = 1,
                 gcmvn: str = "", eps: float = 1e-5) -> None:
        super().__init__()
        self.gmean, self.gstd = None, None
        if gcmvn:
            try:
                if gcmvn.split(".")[-1] == "ark":
                    try:
                        from kaldi_python_io.functional import read_kaldi_mat
                    except ImportError as e:
                        raise RuntimeError("reading a Kaldi .ark cmvn file needs kaldi_python_io") from e
                    stats = th.tensor(read_kaldi_mat(gcmvn), dtype=th.float32)
                    count = stats[0, -1]
                    mean = stats[0, :-1] / count
                    std = (stats[1, :-1] / count - mean**2)**0.5
                else:
                    stats = th.load(gcmvn)
                    mean, std = stats[0], stats[1]
            except FileNotFoundError:
                warnings.warn(f"{gcmvn} not found (no impact when will load checkpoint later) ...")
                mean, std = th.zeros(dim), th.ones(dim)
            self.gmean = nn.Parameter(mean, requires_grad=False)
            self.gstd = nn.Parameter(std, requires_grad=False)
        self.norm_mean, self.norm_var, self.per_band = norm_mean, norm_var, per_band
        self.gcmvn, self.eps = gcmvn, eps

    def extra_repr(self) -> str:
        return (f"norm_mean={self.norm_mean}, norm_var={self.norm_var}, per_band={self.per_band}, "
                f"gcmvn_stats={self.gcmvn}, eps={self.eps:.3e}")

    def global_stats(self, dev: th.device, dim: int):
        """(gmean, gstd) as the kernel reads them: float32, contiguous, on `dev`, `dim` values each.  Statistics loaded
        with `th.load` keep their saved dtype (the reference works in float64 through type promotion, asr.py:576-585),
        so a converted copy is cached and rebuilt when the parameters change; a length that does not match the
        feature dimension raises the reference's broadcast error instead of reading out of bounds."""
        key = (self.gmean._version, self.gstd._version, self.gmean.data_ptr(), self.gstd.data_ptr(), str(dev))
        hit = getattr(self, "_gstats", None)
        if hit is None or hit[0] != key:
            for name, t in (("gmean", self.gmean), ("gstd", self.gstd)):
                if t.device != dev:
                    raise RuntimeError(f"cmvn {name} lives on {t.device}, input on {dev}: move the module first")
            m = self.gmean.detach().reshape(-1).to(th.float32).contiguous()
            v = self.gstd.detach().reshape(-1).to(th.float32).contiguous()
            hit = self._gstats = (key, m, v)
        _, m, v = hit
        for name, t in (("gmean", m), ("gstd", v)):
            if t.numel() != dim:
                raise RuntimeError(f"The size of tensor a ({dim}) must match the size of tensor b ({t.numel()}) at "
                                   f"non-singleton dimension 2 (cmvn {name})")
        return m, v

    def dim_scale(self) -> int:
        return 1

    def forward(self, feats: th.Tensor) -> th.Tensor:
        if not self.norm_mean and not self.norm_var:
            return feats
        if self.gmean is not None:
            if self.norm_mean:
                feats = feats - self.gmean
            return feats / self.gstd if self.norm_var else feats
        axes = -1 if self.per_band else (-1, -2)
        if self.norm_mean:
            feats = feats - th.mean(feats, axes, keepdim=True)
        if self.norm_var:
            var = th.mean(feats**2, axes, keepdim=True) if self.norm_mean else th.var(
                feats, axes, unbiased=False, keepdim=True)
            feats = feats / th.sqrt(var + self.eps)
        return feats


def _host_mask(shape: Tuple[int, int], max_steps: int, num_masks: int, axis: int) -> np.ndarray:
    """One 0/1 mask drawn with Python's `random` in the reference's call order (augment.py:57-82)."""
    m = np.ones(shape, dtype=np.float32)
    L = shape[axis]
    for _ in range(num_masks):
        dur = random.randint(1, max_steps - 1)
        if L - dur <= 0:
            continue
        beg = random.randint(0, L - dur - 1)
        if axis == 1:
            m[:, beg:beg + dur] = 0
        else:
            m[beg:beg + dur, :] = 0
    return m


def tf_mask(batch: int, shape: Tuple[int, int], pm: float = 0.0, ps: float = 0.0, max_bands: int = 30,
            max_frame: int = 40, num_freq_masks: int = 2, num_time_masks: int = 2, device="cpu") -> th.Tensor:
    """Batch of SpecAugment masks N x T x F: host RNG (Q10), one H2D copy (augment.py:13-53)."""
    T, F = shape
    max_bands = min(max_bands, F)
    if ps > 0:
        max_frame = min(max_frame, int(T * ps))
    if pm > 0:
        num_time_masks = min(num_time_masks, int(T * pm))
    masks = np.empty((batch, T, F), dtype=np.float32)
    for n in range(batch):
        fm = _host_mask((T, F), max_bands, num_freq_masks, 1)
        tm = _host_mask((T, F), max_frame, num_time_masks, 0)
        masks[n] = fm * tm
    return th.from_numpy(masks).to(device)


class SpecAugTransform(_Layer):
    """SpecAugment, training only (asr.py:621-684)."""
    _exportable = False

    def __init__(self, p: float = 0.5, adaptive_args: Tuple[float] = (0.0, 0.0), time_args: Tuple[int] = (40, 1),
                 freq_args: Tuple[int] = (30, 1), mask_zero: bool = True) -> None:
        super().__init__()
        assert len(freq_args) == 2 and len(time_args) == 2
        self.fnum, self.tnum = freq_args[1], time_args[1]
        self.mask_zero = mask_zero
        self.F, self.T = freq_args[0], time_args[0]
        self.p = p
        self.pm, self.ps = adaptive_args

    def extra_repr(self) -> str:
        return (f"max_bands={self.F}, max_frame={self.T}, p={self.p}, pm={self.pm}, ps={self.ps}, "
                f"mask_zero={self.mask_zero}, num_freq_masks={self.fnum}, num_time_masks={self.tnum}")

    def forward(self, x: th.Tensor) -> th.Tensor:
        if self.training and th.rand(1).item() < self.p:
            N, T, F = (x.shape[0], x.shape[2], x.shape[3]) if x.dim() == 4 else x.shape
            mask = tf_mask(N, (T, F), pm=self.pm, ps=self.ps, max_bands=self.F, max_frame=self.T,
                           num_freq_masks=self.fnum, num_time_masks=self.tnum, device=x.device)
            if x.is_cuda:
                from .. import ops
                return ops.specaug_apply(x, mask, self.mask_zero).view(x.shape)
            if x.dim() == 4:
                mask = mask.unsqueeze(1)
            x = x * mask if self.mask_zero else th.masked_fill(x, mask == 0, x.mean())
        return x

    def draw_mask(self, N: int, T: int, F: int, device) -> Optional[th.Tensor]:
        """The mask `forward` would apply (same RNG calls in the same order), or None when this call does not augment:
        lets the fused feature kernel multiply it in its epilogue (run_chain)."""
        if self.training and th.rand(1).item() < self.p:
            return tf_mask(N, (T, F), pm=self.pm, ps=self.ps, max_bands=self.F, max_frame=self.T,
                           num_freq_masks=self.fnum, num_time_masks=self.tnum, device=device)
        return None


def splice_feature(feats: th.Tensor, lctx: int = 1, rctx: int = 1, op: str = "cat") -> th.Tensor:
    """Context splicing with edge clamping (utils.py:193-224)."""
    if lctx + rctx == 0:
        return feats
    if op not in ("cat", "stack"):
        raise ValueError(f"Unknown op for feature splicing: {op}")
    T = feats.shape[-2]
    base = th.arange(T, device=feats.device)
    parts = [feats.index_select(-2, (base + c).clamp_(0, T - 1)) for c in range(-lctx, rctx + 1)]
    return th.cat(parts, -1) if op == "cat" else th.stack(parts, -1)


class SpliceTransform(_Layer):
    """Splice + frame subsampling (asr.py:687-728; "next" row f3)."""

    def __init__(self, lctx: int = 0, rctx: int = 0, subsampling_factor: int = 1) -> None:
        super().__init__()
        self.subsampling_factor = subsampling_factor
        self.lctx, self.rctx = max(lctx, 0), max(rctx, 0)

    def extra_repr(self) -> str:
        return f"context=({self.lctx}, {self.rctx}), subsampling_factor={self.subsampling_factor}"

    def dim_scale(self) -> int:
        return 1 + self.rctx + self.lctx

    def forward(self, feats: th.Tensor) -> th.Tensor:
        if feats.is_cuda:
            if self.lctx + self.rctx == 0 and self.subsampling_factor == 1:
                return feats
            from .. import ops
            return ops.splice(feats, self.lctx, self.rctx, self.subsampling_factor)
        feats = splice_feature(feats, lctx=self.lctx, rctx=self.rctx)
        if self.subsampling_factor != 1:
            end = (feats.shape[-2] // self.subsampling_factor) * self.subsampling_factor
            feats = feats[..., :end:self.subsampling_factor, :]
        return feats


class DeltaTransform(_Layer):
    """Delta / delta-delta features (asr.py:731-781; "next" row f3)."""

    def __init__(self, ctx: int = 2, order: int = 2, delta_as_channel: bool = False) -> None:
        super().__init__()
        self.ctx, self.order = ctx, order
        taps = th.arange(-ctx, ctx + 1, dtype=th.float32)
        self.scale = nn.Parameter(taps / sum(i * i for i in range(-ctx, ctx + 1)), requires_grad=False)
        self.delta_as_channel = delta_as_channel

    def extra_repr(self) -> str:
        return f"context={self.ctx}, order={self.order}, delta_as_channel={self.delta_as_channel}"

    def dim_scale(self) -> int:
        return self.order

    def forward(self, feats: th.Tensor) -> th.Tensor:
        if feats.is_cuda and (not self.delta_as_channel or feats.dim() == 3):
            from .. import ops
            return ops.delta(feats, self.scale, self.order, as_channel=self.delta_as_channel)
        outs = [feats]
        for _ in range(self.order):
            ctx = splice_feature(outs[-1], lctx=self.ctx, rctx=self.ctx, op="stack")
            outs.append(th.sum(ctx * self.scale, -1))
        return th.stack(outs, 1) if self.delta_as_channel else th.cat(outs, -1)


# ------------------------------------------------------------------------------------ fused execution
def _feat_desc(power: float, mel: Optional[MelTransform], log: Optional[LogTransform],
               cmvn: Optional[CmvnTransform], dev: th.device, feat_dim: int):
    """Build the `aps_b200_feat_desc` for a [Power][Mel][Log][Cmvn] tail; returns (desc, allband_cmvn).
    `feat_dim`: width of the feature rows the tail produces (num_mels, or the number of STFT bins)."""
    d = _lib.FeatDesc()
    d.power = int(power)
    keep = []
    if mel is not None:
        if mel.filters.device != dev:
            raise RuntimeError(f"mel filters live on {mel.filters.device}, input on {dev}: move the module first")
        start, length, weight, stride = mel.bands(dev)
        d.num_mels, d.mel_stride = mel.num_mels, stride
        d.mel_start, d.mel_len, d.mel_weight = start.data_ptr(), length.data_ptr(), weight.data_ptr()
        keep += [start, length, weight]
    if log is not None:
        d.log_mode = 2 if log.lower_bound > 0 else 1
        d.log_eps, d.log_lower_bound = float(log.eps), float(log.lower_bound)
    allband = None
    if cmvn is not None and (cmvn.norm_mean or cmvn.norm_var):
        d.norm_mean, d.norm_var, d.cmvn_eps = int(cmvn.norm_mean), int(cmvn.norm_var), float(cmvn.eps)
        if cmvn.gmean is not None:
            d.cmvn_mode = 2
            gm, gs = cmvn.global_stats(dev, feat_dim)
            d.gmean, d.gstd = gm.data_ptr(), gs.data_ptr()
            keep += [gm, gs]
        elif cmvn.per_band:
            d.cmvn_mode = 1
        else:
            allband = cmvn            # utterance statistics over (T, F): second tiny kernel
    d._keep = keep
    d.nan_count = 0
    return d, allband


def _match_tail(layers: List[nn.Module], i: int):
    """layers[i:] must start with Magnitude(-1) TFTranspose Power(1|2); then optional Mel, Log, Cmvn.
    Returns (power, mel, log, cmvn, mag_eps, next_index) or None."""
    if i + 2 >= len(layers):
        return None
    mag, tr, pw = layers[i], layers[i + 1], layers[i + 2]
    if not (isinstance(mag, MagnitudeTransform) and mag.dim == -1 and isinstance(tr, TFTransposeTransform)
            and isinstance(pw, PowerTransform) and pw.power in (1, 2)):
        return None
    j = i + 3
    mel = log = cmvn = None
    if j < len(layers) and isinstance(layers[j], MelTransform):
        mel, j = layers[j], j + 1
    if j < len(layers) and isinstance(layers[j], LogTransform):
        log, j = layers[j], j + 1
    if j < len(layers) and isinstance(layers[j], CmvnTransform):
        cmvn, j = layers[j], j + 1
    return pw.power, mel, log, cmvn, float(mag.eps), j


def _apply_allband(out: th.Tensor, cmvn: CmvnTransform, dev: th.device) -> None:
    rows = out.numel() // (out.shape[-1] * out.shape[-2])
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_cmvn_allband(out.data_ptr(), rows, out.shape[-2], out.shape[-1],
                                                     int(cmvn.norm_mean), int(cmvn.norm_var), float(cmvn.eps),
                                                     _lib.stream_ptr(dev)))


def fused_wave_features(spec: SpectrogramTransform, wav: th.Tensor, tail, rescale: bool,
                        utt_preemph: float, nan_count: Optional[th.Tensor] = None,
                        aug: Optional["SpecAugTransform"] = None) -> th.Tensor:
    """F1: waveform N x (C) x S -> features N x (C) x T x D in one kernel.  `nan_count` (int32[1] on the
    device, pre-zeroed) receives the number of NaN values written.  `aug`: a zero-fill SpecAugment layer that follows
    the chain — its mask (host RNG, drawn here exactly as its own forward would) is multiplied in the kernel's epilogue."""
    power, mel, log, cmvn, mag_eps, _ = tail
    if wav.dim() not in (2, 3):
        raise RuntimeError(f"STFT expect 2D/3D tensor, but got {wav.dim():d}D")
    dev = _lib.require_cuda(wav, "the feature transform input")
    x = wav.detach()
    if x.dtype != th.float32:
        x = x.float()
    S = x.shape[-1]
    rows = x.numel() // S
    x = x.reshape(rows, S)
    if x.stride(-1) != 1:
        x = x.contiguous()
    sd = spec.stft_desc(dev, rescale=rescale, utt_preemph=utt_preemph)
    D = mel.num_mels if mel is not None else spec.num_bins
    fd, allband = _feat_desc(power, mel, log, cmvn, dev, D)
    if nan_count is not None and allband is None:
        fd.nan_count = nan_count.data_ptr()
    lib = _lib.load()
    T = lib.aps_b200_num_frames(S, sd.frame_width, sd.hop, sd.center_pad)
    if T < 1:
        raise RuntimeError(f"STFT: {S} samples are too few for one frame of {sd.frame_width}")
    out = th.empty((rows, T, D), dtype=th.float32, device=dev)
    if aug is not None:
        mask = aug.draw_mask(rows, T, D, dev)
        if mask is not None:
            mask = mask.contiguous()
            fd.aug_mask = mask.data_ptr()
            fd._keep.append(mask)
    with th.cuda.device(dev):
        _lib.check(lib.aps_b200_feats_fwd(x.data_ptr(), rows, S, x.stride(0), sd, fd, out.data_ptr(),
                                          _lib.stream_ptr(dev)))
    if allband is not None:
        _apply_allband(out, allband, dev)
    if spec.mode == "torch" and wav.dim() == 3:
        return out.view(rows, 1, T, D)          # same [N*C, 1, ...] quirk as the torch-mode STFT
    return out.view(*wav.shape[:-1], T, D)


def fused_spec_features(packed: th.Tensor, ref_channel: int, tail, extra_cols: int = 0):
    """F1b: packed STFT N x (C) x F x T x 2 -> N x T x (D + extra_cols); returns (out, D)."""
    power, mel, log, cmvn, mag_eps, _ = tail
    dev = _lib.require_cuda(packed, "the packed STFT")
    x = packed.detach()
    if x.dtype != th.float32:
        x = x.float()
    x = x.contiguous()
    if x.dim() == 4:
        N, F, T, _ = x.shape
        C, ref = 1, 0
    elif x.dim() == 5:
        N, C, F, T, _ = x.shape
        ref = ref_channel
    else:
        raise RuntimeError(f"expect a packed STFT N x (C) x F x T x 2, got {x.dim()}D")
    D = mel.num_mels if mel is not None else F
    if mel is not None and mel.filters.shape[-1] != F:
        raise RuntimeError(f"mel filters expect {mel.filters.shape[-1]} bins, the packed STFT has {F}")
    fd, allband = _feat_desc(power, mel, log, cmvn, dev, D)
    out = th.empty((N, T, D + extra_cols), dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_spec_feats_fwd(x.data_ptr(), N, C, ref, F, T, mag_eps, fd, out.data_ptr(),
                                                       D + extra_cols, _lib.stream_ptr(dev)))
    if allband is not None:
        if extra_cols:
            raise RuntimeError("all-band cmvn cannot be fused with extra feature columns")
        _apply_allband(out, allband, dev)
    return out, D


def run_chain(layers: List[nn.Module], x: th.Tensor, nan_count: Optional[th.Tensor] = None) -> th.Tensor:
    """Run a FeatureTransform layer list on a waveform, fusing the spectral run into F1."""
    i, n = 0, len(layers)
    rescale, utt_pre = False, 0.0
    while i < n:
        lay = layers[i]
        # look ahead: [Rescale] [PreEmphasis] Spectrogram + tail  ==> one kernel
        j = i
        r, e = False, 0.0
        emph_layer = None
        if j < n and isinstance(layers[j], RescaleTransform) and layers[j].rescale == MAX_INT16 * 1.0:
            r, j = True, j + 1
        if j < n and isinstance(layers[j], SpeedPerturbTransform) and not layers[j].training:
            layers[j](x)                 # identity in eval mode; resets its `last_choice`
            j += 1
        if j < n and isinstance(layers[j], PreEmphasisTransform):
            emph_layer, e, j = layers[j], float(layers[j].pre_emphasis), j + 1
        if j < n and isinstance(layers[j], SpectrogramTransform) and x.is_cuda:
            tail = _match_tail(layers, j + 1)
            if tail is not None:
                nxt = tail[5]
                cm = tail[3]
                allband = cm is not None and cm.gmean is None and not cm.per_band and (cm.norm_mean or cm.norm_var)
                aug = None
                if (nxt < n and isinstance(layers[nxt], SpecAugTransform) and layers[nxt].mask_zero and x.dim() == 2
                        and not allband):
                    aug, nxt = layers[nxt], nxt + 1      # SpecAugment (zero fill) rides in the kernel's epilogue
                y = fused_wave_features(layers[j], x, tail, rescale=r, utt_preemph=max(e, 0.0), nan_count=nan_count,
                                        aug=aug)
                if emph_layer is not None and e > 0 and not r:
                    emph_layer(x)        # keep the reference's in-place side effect on the caller's wav (Q9)
                x, i = y, nxt
                continue
        x = lay(x)
        i += 1
    return x


# ------------------------------------------------------------------------------------ the transform
class FeatureTransform(nn.Module):
    """Feature transform for ASR tasks — see the reference docstring (asr.py:785-836) for the arguments;
    names, defaults and the `feats` token grammar are identical."""

    def __init__(self,
                 feats: str = "fbank-log-cmvn",
                 frame_len: int = 400,
                 frame_hop: int = 160,
                 window: str = "hamm",
                 center: bool = False,
                 round_pow_of_two: bool = True,
                 stft_normalized: bool = False,
                 stft_mode: str = "librosa",
                 audio_norm: bool = True,
                 pre_emphasis: float = 0.97,
                 use_power: bool = False,
                 sr: int = 16000,
                 speed_perturb: str = "0.9,1.0,1.1",
                 log_lower_bound: float = 0,
                 num_mels: int = 80,
                 mel_matrix: str = "",
                 mel_coeff_norm: bool = False,
                 min_freq: int = 0,
                 max_freq: Optional[int] = None,
                 num_ceps: int = 13,
                 lifter: float = 0,
                 aug_prob: float = 0,
                 aug_adaptive_args: Tuple[float] = (0, 0),
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
                 eps: float = EPSILON) -> None:
        super().__init__()
        if not feats:
            raise ValueError("FeatureTransform: 'feats' can not be empty")
        stft_kw = dict(mode=stft_mode, window=window, center=center, normalized=stft_normalized,
                       pre_emphasis=pre_emphasis, round_pow_of_two=round_pow_of_two)
        mel_kw = dict(round_pow_of_two=round_pow_of_two, sr=sr, fmin=min_freq, fmax=max_freq, num_mels=num_mels,
                      coeff_norm=mel_coeff_norm, mel_matrix=mel_matrix, requires_grad=requires_grad)

        def spectral(with_mel: bool) -> List[nn.Module]:
            block = [SpectrogramTransform(frame_len, frame_hop, **stft_kw), MagnitudeTransform(dim=-1),
                     TFTransposeTransform(), PowerTransform(power=2 if use_power else 1)]
            if with_mel:
                block.append(MelTransform(frame_len, **mel_kw))
            return block

        layers: List[nn.Module] = [] if audio_norm else [RescaleTransform()]
        dim = 0
        self.spectra_index = -1
        self.perturb_index = -1
        for tok in feats.split("-"):
            if tok == "perturb":
                self.perturb_index = len(layers)
                layers.append(SpeedPerturbTransform(sr=sr, perturb=speed_perturb))
            elif tok == "emph":
                layers.append(PreEmphasisTransform(pre_emphasis=pre_emphasis))
            elif tok in ("spectrogram", "fbank", "mfcc"):
                self.spectra_index = len(layers)
                layers += spectral(tok != "spectrogram")
                if tok == "mfcc":
                    layers += [LogTransform(eps=eps, lower_bound=log_lower_bound),
                               DiscreteCosineTransform(num_ceps=num_ceps, num_mels=num_mels, lifter=lifter)]
                dim = layers[-1].dim() if tok != "spectrogram" else layers[self.spectra_index].dim()
            elif tok == "trans":
                layers.append(TFTransposeTransform())
            elif tok == "pow":
                layers.append(PowerTransform())
            elif tok == "mel":
                layers.append(MelTransform(frame_len, **mel_kw))
                dim = layers[-1].dim()
            elif tok == "log":
                layers.append(LogTransform(eps=eps, lower_bound=log_lower_bound))
            elif tok == "abs":
                layers.append(AbsTransform(eps=eps))
            elif tok == "dct":
                layers.append(DiscreteCosineTransform(num_ceps=num_ceps, num_mels=num_mels, lifter=lifter))
                dim = layers[-1].dim()
            elif tok == "cmvn":
                layers.append(CmvnTransform(norm_mean=norm_mean, norm_var=norm_var, per_band=norm_per_band,
                                            gcmvn=gcmvn, dim=dim, eps=eps))
            elif tok == "aug":
                layers.append(SpecAugTransform(p=aug_prob, adaptive_args=aug_adaptive_args,
                                               freq_args=aug_freq_args, time_args=aug_time_args,
                                               mask_zero=aug_mask_zero))
            elif tok == "splice":
                layers.append(SpliceTransform(lctx=lctx, rctx=rctx, subsampling_factor=subsampling_factor))
                dim *= (1 + lctx + rctx)
            elif tok == "delta":
                layers.append(DeltaTransform(ctx=delta_ctx, order=delta_order,
                                             delta_as_channel=delta_as_channel))
                dim *= (1 + delta_order)
            else:
                raise RuntimeError(f"Unknown token {tok} in {feats}")
        self.transform = nn.Sequential(*layers)
        self.feats_dim = dim
        self.subsampling_factor = subsampling_factor

    def dim(self) -> int:
        return self.feats_dim

    def num_frames(self, inp_len: Optional[th.Tensor]) -> Optional[th.Tensor]:
        """Number of frames per utterance — integer exact (asr.py:1003-1019)."""
        if inp_len is None:
            return None
        if self.spectra_index == -1:
            warnings.warn("SpectrogramTransform layer is not found, return input as the #num_frames")
            return inp_len
        if self.perturb_index != -1:
            inp_len = self.transform[self.perturb_index].output_length(inp_len)
        frames = self.transform[self.spectra_index].num_frames(inp_len)
        return th.div(frames, self.subsampling_factor, rounding_mode="trunc")

    def forward(self, inp_pad: th.Tensor, inp_len: Optional[th.Tensor]) -> AsrReturnType:
        """inp_pad: N x (C) x S waveform (or features for token strings without a spectral token);
        returns (feats N x (C) x T x D, num_frames or None) — asr.py:1021-1033."""
        if isinstance(inp_pad, th.Tensor) and not inp_pad.is_cuda and self.spectra_index != -1:
            _lib.require_cuda(inp_pad, "AsrTransform input")
        counter = None
        if isinstance(inp_pad, th.Tensor) and inp_pad.is_cuda and self._nan_by_kernel():
            counter = th.zeros(1, dtype=th.int32, device=inp_pad.device)
        feats = run_chain(list(self.transform), inp_pad, nan_count=counter)
        return check_valid(feats, self.num_frames(inp_len), counter)

    def _nan_by_kernel(self) -> bool:
        """True when the whole chain after the waveform layers is the fused run (so the kernel's NaN
        counter covers the final features; linear tails such as splice/delta/aug cannot create NaNs
        but all-band CMVN runs as a second kernel and is re-checked on the output)."""
        if self.spectra_index == -1:
            return False
        layers = list(self.transform)
        tail = _match_tail(layers, self.spectra_index + 1)
        if tail is None:
            return False
        cm = tail[3]
        if cm is not None and cm.gmean is None and not cm.per_band and (cm.norm_mean or cm.norm_var):
            return False
        return all(isinstance(m, (SpecAugTransform, SpliceTransform, DeltaTransform)) for m in layers[tail[5]:])
