"""Oracle (CPU, fp32) for rows a1–a13 of SURVEY.md §8: window, DFT kernel, frame
counts, dense-DFT STFT / iSTFT, |X|, mel, log, CMVN, SpecAug application.

TEST INFRASTRUCTURE — see oracle/__init__.py.  All citations are into
/root/reference.  The arithmetic deliberately follows the reference's *algorithm*
(a dense DFT done as a GEMM over unfolded frames, a two-sided spectrum that is
sliced afterwards, overlap-add through a transposed convolution) because this
module is also what `bench.py` times as the reference CPU path.
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import torch as th
import torch.nn.functional as F

F32_EPS = float(np.finfo(np.float32).eps)  # aps/const.py:17  EPSILON
WINDOWS = ("bartlett", "hann", "hamm", "blackman", "rect", "sqrthann")


# ----------------------------------------------------------------------------- a1
def window(name: str, frame_len: int) -> th.Tensor:
    """Periodic analysis window.  Ref: aps/transform/utils.py:30-59 (init_window)."""
    if name not in WINDOWS:
        raise RuntimeError(f"Unknown window type: {name}")
    if name == "rect":
        return th.ones(frame_len)
    maker = {
        "hann": th.hann_window,
        "sqrthann": th.hann_window,
        "hamm": th.hamming_window,
        "blackman": th.blackman_window,
        "bartlett": th.bartlett_window,
    }[name]
    coeff = maker(frame_len, periodic=True)
    return coeff**0.5 if name == "sqrthann" else coeff


def fft_size_of(frame_len: int, round_pow_of_two: bool = True, mode: str = "librosa") -> int:
    """Ref: utils.py:83-86 (kaldi mode always rounds up)."""
    if round_pow_of_two or mode == "kaldi":
        return 2**math.ceil(math.log2(frame_len))
    return frame_len


# ----------------------------------------------------------------------------- a2
def dft_kernel(frame_len: int,
               win: th.Tensor,
               round_pow_of_two: bool = True,
               normalized: bool = False,
               inverse: bool = False,
               mode: str = "librosa") -> Tuple[th.Tensor, th.Tensor]:
    """(K[2B,1,W], w[W]): stacked real/imag DFT rows.  Ref: utils.py:62-112 (init_kernel).

    librosa mode centre-pads the window to the FFT size (:88-90); kaldi mode keeps the
    window at frame_len and drops DFT columns >= frame_len (:103-104), i.e. the frame
    is zero-padded at its tail.
    """
    if mode not in ("librosa", "kaldi"):
        raise ValueError(f"Unsupported mode: {mode}")
    B = fft_size_of(frame_len, round_pow_of_two, mode)
    if mode == "librosa" and B != frame_len:
        left = (B - frame_len) // 2
        win = F.pad(win, (left, B - frame_len - left))
    scale = B**0.5 if normalized else 1.0
    spec = th.fft.fft(th.eye(B) / scale, dim=-1)           # [n, k]
    mat = th.stack([spec.real, spec.imag], dim=-1)          # W x B x 2
    if mode == "kaldi":
        mat = mat[:frame_len]
    if inverse and not normalized:
        mat = mat / B
    mat = mat.permute(2, 1, 0).reshape(2 * B, 1, mat.shape[0])
    return mat.contiguous(), win


# ----------------------------------------------------------------------------- a3
def num_frames(wav_len: th.Tensor, win_length: int, hop: int, center: bool) -> th.Tensor:
    """Frame count (integer, must be bit exact).  Ref: utils.py:653-662.

    NOTE the reference adds `win_length` to its argument IN PLACE when center=True
    (Q5 in SURVEY.md); this restatement is pure and returns the value the reference
    returns on its first call.
    """
    assert int(th.sum(wav_len <= win_length)) == 0
    eff = wav_len + win_length if center else wav_len
    return th.div(eff - win_length, hop, rounding_mode="trunc") + 1


# ----------------------------------------------------------------------------- a4
def stft_dense(wav: th.Tensor,
               K: th.Tensor,
               w: th.Tensor,
               hop: int,
               pre_emphasis: float = 0.0,
               onesided: bool = True,
               center: bool = False,
               polar: bool = False,
               eps: float = F32_EPS) -> th.Tensor:
    """Dense-DFT STFT: N x (C) x S -> N x (C) x F x T x 2.  Ref: utils.py:227-290.

    pre_emphasis > 0 takes the unfold+matmul branch with PER-FRAME Kaldi pre-emphasis
    (:263-272); otherwise a strided conv1d (:274).
    """
    if wav.dim() not in (2, 3):
        raise RuntimeError(f"STFT expect 2D/3D tensor, but got {wav.dim():d}D")
    lead = wav.shape[:-1]
    S = wav.shape[-1]
    W = K.shape[-1]
    x = wav.reshape(-1, 1, S)
    if center:
        x = F.pad(x, (W // 2, W // 2), mode="reflect")
    basis = K * w                                           # 2B x 1 x W
    if pre_emphasis > 0:
        fr = F.unfold(x[:, None], (1, W), stride=hop)       # NC x W x T
        head = fr[:, :1] * (1 - pre_emphasis)
        rest = fr[:, 1:] - pre_emphasis * fr[:, :-1]
        fr = th.cat([head, rest], 1)
        packed = th.matmul(basis[:, 0][None], fr)           # NC x 2B x T
    else:
        packed = F.conv1d(x, basis, stride=hop)
    nb = K.shape[0] // 2
    re, im = packed[:, :nb], packed[:, nb:]
    if onesided:
        keep = K.shape[0] // 4 + 1
        re, im = re[:, :keep], im[:, :keep]
    if polar:
        out = th.stack([(re**2 + im**2 + eps)**0.5, th.atan2(im, re)], -1)
    else:
        out = th.stack([re, im], -1)
    return out.reshape(*lead, *out.shape[1:])


# ----------------------------------------------------------------------------- a5
def stft_torch(wav: th.Tensor, frame_len: int, hop: int, w: th.Tensor, n_fft: int,
               normalized: bool = False, onesided: bool = True, center: bool = False,
               polar: bool = False, eps: float = F32_EPS) -> th.Tensor:
    """`stft_mode="torch"`: th.stft wrapper, no pre-emphasis.  Ref: utils.py:363-415."""
    if wav.dim() not in (2, 3):
        raise RuntimeError(f"STFT expect 2D/3D tensor, but got {wav.dim():d}D")
    lead = wav.shape[:-1]
    z = th.stft(wav.reshape(-1, wav.shape[-1]), n_fft, hop_length=hop, win_length=w.shape[-1],
                window=w, center=center, normalized=normalized, onesided=onesided,
                return_complex=True)
    re, im = z.real, z.imag
    if polar:
        out = th.stack([(re**2 + im**2 + eps)**0.5, th.atan2(im, re)], -1)
    else:
        out = th.stack([re, im], -1)
    return out.reshape(*lead, *out.shape[1:])


# ----------------------------------------------------------------------------- a13
def istft_dense(spec: th.Tensor, K: th.Tensor, w: th.Tensor, hop: int, onesided: bool = True,
                center: bool = False, polar: bool = False, eps: float = F32_EPS) -> th.Tensor:
    """Dense-iDFT inverse STFT with window^2 overlap-add de-normalisation:
    (N) x F x T x 2 -> N x S.  Ref: utils.py:293-360 (_inverse_stft)."""
    if spec.dim() == 3:
        spec = spec[None]
    if spec.dim() != 4:
        raise RuntimeError(f"Expect 4D tensor, but got {spec.dim()}D")
    a, b = spec[..., 0], spec[..., 1]
    re, im = (a * th.cos(b), a * th.sin(b)) if polar else (a, b)
    if onesided:
        mirror = list(range(K.shape[0] // 4 - 1, 0, -1))
        re = th.cat([re, re[:, mirror]], 1)
        im = th.cat([im, -im[:, mirror]], 1)
    packed = th.cat([re, im], 1)
    wav = F.conv_transpose1d(packed, K * w, stride=hop)
    T, W = packed.shape[-1], w.shape[0]
    sq = (w**2)[:, None].expand(W, T)[None]
    norm = F.conv_transpose1d(sq, th.eye(W)[:, None], stride=hop)
    if center:
        p = K.shape[-1] // 2
        wav, norm = wav[..., p:-p], norm[..., p:-p]
    return (wav / (norm + eps)).squeeze(1)


def istft_torch(spec: th.Tensor, hop: int, w: th.Tensor, n_fft: int, normalized: bool = False,
                onesided: bool = True, center: bool = False, polar: bool = False) -> th.Tensor:
    """Ref: utils.py:418-469 (_pytorch_istft)."""
    if spec.dim() == 3:
        spec = spec[None]
    if spec.dim() != 4:
        raise RuntimeError(f"Expect 4D tensor, but got {spec.dim()}D")
    if polar:
        spec = th.stack([spec[..., 0] * th.cos(spec[..., 1]), spec[..., 0] * th.sin(spec[..., 1])], -1)
    z = th.view_as_complex(spec.contiguous())
    return th.istft(z, n_fft, hop_length=hop, win_length=w.shape[-1], window=w, center=center,
                    normalized=normalized, onesided=onesided, return_complex=False)


# ----------------------------------------------------------------------------- a7
def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0**(np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(frame_len: int, round_pow_of_two: bool = True, num_bins: Optional[int] = None,
                   sr: int = 16000, num_mels: int = 80, fmin: float = 0.0,
                   fmax: Optional[float] = None, norm: bool = False) -> th.Tensor:
    """HTK-scale triangular filterbank [num_mels, N/2+1], float32.

    Ref: aps/transform/utils.py:115-156 (mel_filter) which delegates the values to
    librosa==0.8.1 `filters.mel(sr, N, n_mels, fmin, fmax, htk=True, norm=None|"slaney")`
    — a third-party dependency absent from /root/reference AND from this image; the
    reference's tests pin only the shape.  PARITY UNPINNED for the values: this is a
    restatement of librosa's published algorithm.
    """
    if num_bins is None:
        N = 2**math.ceil(math.log2(frame_len)) if round_pow_of_two else frame_len
    else:
        N = (num_bins - 1) * 2
    nyq = sr // 2
    fmax = nyq if fmax is None else min(fmax + nyq if fmax < 0 else fmax, nyq)
    fmin = max(0, fmin)
    nb = N // 2 + 1
    centres = np.linspace(0, float(sr) / 2, nb)
    edges = mel_to_hz_htk(np.linspace(hz_to_mel_htk(fmin), hz_to_mel_htk(fmax), num_mels + 2))
    width = np.diff(edges)
    offs = edges[:, None] - centres[None, :]
    fb = np.zeros((num_mels, nb), dtype=np.float32)
    for m in range(num_mels):
        rise = -offs[m] / width[m]
        fall = offs[m + 2] / width[m + 1]
        fb[m] = np.maximum(0, np.minimum(rise, fall))
    if norm:
        fb *= (2.0 / (edges[2:num_mels + 2] - edges[:num_mels]))[:, None].astype(np.float32)
    return th.from_numpy(fb)


# ----------------------------------------------------------------------------- a6, a8, a9
def magnitude(packed: th.Tensor, eps: float = 0.0) -> th.Tensor:
    """sqrt(re^2+im^2+eps) over the trailing axis.  Ref: aps/transform/asr.py:296-303."""
    return th.sqrt(th.sum(packed**2, -1) + eps)


def log_compress(x: th.Tensor, eps: float = F32_EPS, lower_bound: float = 0.0) -> th.Tensor:
    """Ref: asr.py:453-464 (LogTransform)."""
    return th.log(lower_bound + x) if lower_bound > 0 else th.log(th.clamp(x, min=eps))


def cmvn(x: th.Tensor, norm_mean: bool = True, norm_var: bool = True, per_band: bool = True,
         gmean: Optional[th.Tensor] = None, gstd: Optional[th.Tensor] = None,
         eps: float = F32_EPS) -> th.Tensor:
    """Ref: asr.py:576-618.  per_band reduces over the LAST axis (mel, per frame — Q6);
    all_band over the last two axes; global stats override both."""
    if not norm_mean and not norm_var:
        return x
    if gmean is not None:
        if norm_mean:
            x = x - gmean
        if norm_var:
            x = x / gstd
        return x
    dims = -1 if per_band else (-1, -2)
    if norm_mean:
        x = x - th.mean(x, dims, keepdim=True)
    if norm_var:
        var = th.mean(x**2, dims, keepdim=True) if norm_mean else th.var(
            x, dims, unbiased=False, keepdim=True)
        x = x / th.sqrt(var + eps)
    return x


# ----------------------------------------------------------------------------- a10
def specaug_apply(x: th.Tensor, mask: th.Tensor, mask_zero: bool = True) -> th.Tensor:
    """Apply a host-generated 0/1 mask N x T x F.  Ref: asr.py:677-683 (the mask itself is
    drawn on the host by aps/transform/augment.py:13-82 and is NOT restated here)."""
    if x.dim() == 4:
        mask = mask.unsqueeze(1)
    return x * mask if mask_zero else th.masked_fill(x, mask == 0, x.mean())


# ----------------------------------------------------------------------------- a11
@dataclass
class AsrFeatCfg:
    """Subset of aps/transform/asr.py:837-875 that the fused device path covers."""
    feats: str = "fbank-log-cmvn"
    frame_len: int = 400
    frame_hop: int = 160
    window: str = "hamm"
    center: bool = False
    round_pow_of_two: bool = True
    stft_normalized: bool = False
    stft_mode: str = "librosa"
    audio_norm: bool = True
    pre_emphasis: float = 0.97
    use_power: bool = False
    sr: int = 16000
    log_lower_bound: float = 0.0
    num_mels: int = 80
    mel_coeff_norm: bool = False
    min_freq: int = 0
    max_freq: Optional[int] = None
    norm_mean: bool = True
    norm_var: bool = True
    norm_per_band: bool = True
    eps: float = F32_EPS
    gmean: Optional[th.Tensor] = field(default=None, repr=False)
    gstd: Optional[th.Tensor] = field(default=None, repr=False)
    num_ceps: int = 13
    lifter: float = 0
    lctx: int = 1
    rctx: int = 1
    subsampling_factor: int = 1
    delta_ctx: int = 2
    delta_order: int = 2


def dct_matrix(num_ceps: int, num_mels: int) -> th.Tensor:
    """Orthonormal DCT-II rows (scipy.fftpack.dct(eye, norm="ortho")[:, :num_ceps].T, asr.py:483-487): [num_ceps, num_mels]."""
    n = th.arange(num_mels, dtype=th.float64)
    k = th.arange(num_ceps, dtype=th.float64)[:, None]
    m = th.cos(math.pi * (2 * n + 1) * k / (2 * num_mels)) * math.sqrt(2.0 / num_mels)
    m[0] = m[0] / math.sqrt(2.0)
    return m.float()


def splice(feats: th.Tensor, lctx: int, rctx: int, op: str = "cat") -> th.Tensor:
    """aps/transform/utils.py:193-224 (edge frames repeat)."""
    if lctx + rctx == 0:
        return feats
    T = feats.shape[-2]
    ctx = [feats.index_select(-2, th.arange(c, c + T).clamp(0, T - 1)) for c in range(-lctx, rctx + 1)]
    return th.cat(ctx, -1) if op == "cat" else th.stack(ctx, -1)


def delta(feats: th.Tensor, ctx: int = 2, order: int = 2) -> th.Tensor:
    """asr.py:731-781 (delta_as_channel=False)."""
    scale = th.arange(-ctx, ctx + 1, dtype=th.float32) / sum(i * i for i in range(-ctx, ctx + 1))
    out = [feats]
    for _ in range(order):
        out.append(th.sum(splice(out[-1], ctx, ctx, "stack") * scale, -1))
    return th.cat(out, -1)


class AsrFeatures:
    """Functional restatement of FeatureTransform for the token chains
    `[emph-](spectrogram|fbank)[-log][-cmvn]`.  Ref: asr.py:876-1033."""

    def __init__(self, cfg: AsrFeatCfg):
        self.cfg = c = cfg
        self.tokens = c.feats.split("-")
        self.w0 = window(c.window, c.frame_len)
        if c.stft_mode == "torch":
            self.K = None
            self.w = self.w0
            self.nfft = fft_size_of(c.frame_len, c.round_pow_of_two)
            self.win_length = self.nfft
            self.pre_emphasis = 0.0                          # utils.py:643
        else:
            self.K, self.w = dft_kernel(c.frame_len, self.w0, c.round_pow_of_two,
                                        c.stft_normalized, False, c.stft_mode)
            self.nfft = self.K.shape[0] // 2
            self.win_length = self.K.shape[2]
            self.pre_emphasis = c.pre_emphasis
        self.num_bins = self.nfft // 2 + 1
        self.mel = mel_filterbank(c.frame_len, c.round_pow_of_two, None, c.sr, c.num_mels,
                                  c.min_freq, c.max_freq, c.mel_coeff_norm)

    def stft(self, wav: th.Tensor) -> th.Tensor:
        c = self.cfg
        if c.stft_mode == "torch":
            return stft_torch(wav, c.frame_len, c.frame_hop, self.w, self.nfft,
                              c.stft_normalized, True, c.center)
        return stft_dense(wav, self.K, self.w, c.frame_hop, self.pre_emphasis, True, c.center)

    def frames(self, lens: Optional[th.Tensor]) -> Optional[th.Tensor]:
        if lens is None:
            return None
        return num_frames(lens.clone(), self.win_length, self.cfg.frame_hop, self.cfg.center)

    def __call__(self, wav: th.Tensor, lens: Optional[th.Tensor] = None):
        c = self.cfg
        x = wav
        if not c.audio_norm:
            x = th.round(x * 32767.0)                        # asr.py:84,880 (RescaleTransform)
        for tok in self.tokens:
            if tok == "emph":                                # asr.py:111-113 (utterance level)
                if c.pre_emphasis > 0:
                    x = th.cat([x[..., :1], x[..., 1:] - c.pre_emphasis * x[..., :-1]], -1)
            elif tok in ("spectrogram", "fbank", "mfcc"):
                x = magnitude(self.stft(x)).transpose(-1, -2)
                x = x**(2 if c.use_power else 1)             # asr.py:357
                if tok != "spectrogram":
                    x = F.linear(x, self.mel)                # asr.py:427
                if tok == "mfcc":                            # asr.py:931-945: fbank -> log -> DCT (+ lifter)
                    x = F.linear(log_compress(x, c.eps, c.log_lower_bound), dct_matrix(c.num_ceps, c.num_mels))
                    if c.lifter > 0:
                        x = x * (1 + c.lifter * 0.5 * th.sin(math.pi * th.arange(1, 1 + c.num_ceps) / c.lifter))
            elif tok == "splice":                            # asr.py:687-728
                x = splice(x, max(c.lctx, 0), max(c.rctx, 0))
                if c.subsampling_factor != 1:
                    end = (x.shape[-2] // c.subsampling_factor) * c.subsampling_factor
                    x = x[..., :end:c.subsampling_factor, :]
            elif tok == "delta":
                x = delta(x, c.delta_ctx, c.delta_order)
            elif tok == "log":
                x = log_compress(x, c.eps, c.log_lower_bound)
            elif tok == "cmvn":
                x = cmvn(x, c.norm_mean, c.norm_var, c.norm_per_band, c.gmean, c.gstd, c.eps)
            else:
                raise RuntimeError(f"oracle does not cover token {tok}")
        nf = self.frames(lens)
        if nf is not None:                                   # asr.py:46-53 (check_valid)
            x = x[..., :int(nf.max()), :]
        return x, nf
