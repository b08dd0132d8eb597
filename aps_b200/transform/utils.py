"""STFT / iSTFT layers and their helpers with the surface of the reference's
`aps/transform/utils.py`, backed by the sm_100a kernels F2 / F3 (csrc/frontend.cu, csrc/istft.cu).

Drop-in facts kept on purpose:
  * `K` (the stacked DFT matrix) and `w` stay `nn.Parameter(requires_grad=False)` with the reference's
    shapes so checkpoints load strictly (utils.py:631-632) — the kernels never read `K`: they run an
    FFT, `K` is state-dict ballast only; `w` IS the window the kernels use;
  * `num_frames` mutates its argument in place when `center=True` (utils.py:658-659, Q5);
  * `mode="torch"` silently drops pre-emphasis (utils.py:643, Q4).
"""
import math
from typing import Optional, Tuple

import numpy as np
import torch as th
import torch.nn as nn

from .. import _lib

EPSILON = float(np.finfo(np.float32).eps)   # aps/const.py:17
MAX_INT16 = 32767                           # aps/const.py:18
_WINDOWS = ("bartlett", "hann", "hamm", "blackman", "rect", "sqrthann")


def export_jit(transform: nn.Module) -> nn.Module:
    """Keep the exportable layers only (utils.py:22-27)."""
    return nn.Sequential(*[m for m in transform if m.exportable()])


def init_window(wnd: str, frame_len: int, device="cpu") -> th.Tensor:
    """Periodic window coefficients (utils.py:30-59)."""
    if wnd not in _WINDOWS:
        raise RuntimeError(f"Unknown window type: {wnd}")
    if wnd == "rect":
        return th.ones(frame_len, device=device)
    base = {"hann": th.hann_window, "sqrthann": th.hann_window, "hamm": th.hamming_window,
            "blackman": th.blackman_window, "bartlett": th.bartlett_window}[wnd]
    c = base(frame_len, periodic=True)          # periodic: matches librosa (utils.py:54-56)
    if wnd == "sqrthann":
        c = c**0.5
    return c.to(device)


def fft_size(frame_len: int, round_pow_of_two: bool = True, mode: str = "librosa") -> int:
    if round_pow_of_two or mode == "kaldi":
        return 2**math.ceil(math.log2(frame_len))
    return frame_len


def init_kernel(frame_len: int,
                frame_hop: int,
                window: th.Tensor,
                round_pow_of_two: bool = True,
                normalized: bool = False,
                inverse: bool = False,
                mode: str = "librosa") -> Tuple[th.Tensor, th.Tensor]:
    """(K, w) with the reference's shapes: K [2B, 1, W] = real rows then imaginary rows of the
    (scaled) DFT matrix, w [W] the (centre-padded) window (utils.py:62-112)."""
    if mode not in ("librosa", "kaldi"):
        raise ValueError(f"Unsupported mode: {mode}")
    B = fft_size(frame_len, round_pow_of_two, mode)
    if mode == "librosa" and B != frame_len:
        lpad = (B - frame_len) // 2
        window = th.cat([window.new_zeros(lpad), window, window.new_zeros(B - frame_len - lpad)])
    W = frame_len if mode == "kaldi" else B
    k = th.arange(B, dtype=th.float64)[:, None]
    n = th.arange(W, dtype=th.float64)[None, :]
    ang = 2.0 * math.pi * ((k * n) % B) / B
    scale = 1.0
    if normalized:
        scale = B**-0.5
    elif inverse:
        scale = 1.0 / B
    K = th.cat([th.cos(ang), -th.sin(ang)], 0) * scale
    return K.to(th.float32).reshape(2 * B, 1, W).to(window.device), window


def _hz2mel(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def _mel2hz(m):
    return 700.0 * (10.0**(np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filter(frame_len: int,
               round_pow_of_two: bool = True,
               num_bins: Optional[int] = None,
               sr: int = 16000,
               num_mels: int = 80,
               fmin: float = 0.0,
               fmax: Optional[float] = None,
               norm: bool = False) -> th.Tensor:
    """HTK mel filterbank [num_mels, N/2+1] (utils.py:115-156; values defined by librosa 0.8.1
    `filters.mel(htk=True, norm=None|"slaney")`, which is not importable here — this follows the
    published construction: linear-in-mel band edges, triangles max(0, min(rise, fall)))."""
    if num_bins is None:
        N = 2**math.ceil(math.log2(frame_len)) if round_pow_of_two else frame_len
    else:
        N = (num_bins - 1) * 2
    upper = sr // 2
    fmax = upper if fmax is None else min(fmax + upper if fmax < 0 else fmax, upper)
    fmin = max(0, fmin)
    F = N // 2 + 1
    bins = np.linspace(0, float(sr) / 2, F)
    edge = _mel2hz(np.linspace(_hz2mel(fmin), _hz2mel(fmax), num_mels + 2))
    step = np.diff(edge)
    fb = np.zeros((num_mels, F), dtype=np.float32)
    for m in range(num_mels):
        up = (bins - edge[m]) / step[m]
        down = (edge[m + 2] - bins) / step[m + 1]
        fb[m] = np.maximum(0.0, np.minimum(up, down))
    if norm:
        fb *= (2.0 / (edge[2:] - edge[:-2]))[:, None].astype(np.float32)
    return th.from_numpy(fb)


# ------------------------------------------------------------------------------------------------
class STFTBase(nn.Module):
    """Common state of STFT / iSTFT (utils.py:594-675)."""

    def __init__(self,
                 frame_len: int,
                 frame_hop: int,
                 window: str = "sqrthann",
                 round_pow_of_two: bool = True,
                 normalized: bool = False,
                 pre_emphasis: float = 0,
                 onesided: bool = True,
                 inverse: bool = False,
                 center: bool = False,
                 mode: str = "librosa") -> None:
        super().__init__()
        if mode != "torch":
            K, w = init_kernel(frame_len, frame_hop, init_window(window, frame_len),
                               round_pow_of_two=round_pow_of_two, normalized=normalized,
                               inverse=inverse, mode=mode)
            self.K = nn.Parameter(K, requires_grad=False)
            self.w = nn.Parameter(w, requires_grad=False)
            self.num_bins = self.K.shape[0] // 4 + 1
            self.pre_emphasis = pre_emphasis
            self.win_length = self.K.shape[2]
        else:
            self.K = None
            self.w = nn.Parameter(init_window(window, frame_len), requires_grad=False)
            nfft = fft_size(frame_len, round_pow_of_two)
            self.num_bins = nfft // 2 + 1
            self.pre_emphasis = 0
            self.win_length = nfft
        self.frame_len = frame_len
        self.frame_hop = frame_hop
        self.window = window
        self.normalized = normalized
        self.onesided = onesided
        self.center = center
        self.mode = mode
        self.inverse = inverse
        self._wcache = None

    # -- geometry shared with the kernels ---------------------------------------------------------
    @property
    def nfft(self) -> int:
        return (self.num_bins - 1) * 2

    def center_pad(self) -> int:
        if not self.center:
            return 0
        return self.nfft // 2 if self.mode == "torch" else self.win_length // 2

    def kernel_window(self) -> th.Tensor:
        """Window of `win_length` samples as the kernels want it (torch mode: centre-pad to nfft)."""
        if self.mode != "torch" or self.w.shape[0] == self.nfft:
            return self.w
        key = (self.w.data_ptr(), self.w._version, self.w.device)
        if self._wcache is None or self._wcache[0] != key:
            lpad = (self.nfft - self.w.shape[0]) // 2
            w = th.zeros(self.nfft, dtype=th.float32, device=self.w.device)
            w[lpad:lpad + self.w.shape[0]] = self.w.detach()
            self._wcache = (key, w)
        return self._wcache[1]

    def stft_desc(self, dev: th.device, rescale: bool = False, utt_preemph: float = 0.0) -> "_lib.StftDesc":
        if not self.onesided:
            raise RuntimeError("aps_b200: only onesided=True STFTs are implemented on the B200 path")
        nfft = self.nfft
        if nfft < 64 or nfft > 1024 or nfft & (nfft - 1):
            raise RuntimeError(f"aps_b200: unsupported FFT size {nfft} (need a power of two in [64, 1024]; "
                               "use round_pow_of_two=True)")
        if self.w.device != dev:
            raise RuntimeError(f"STFT window lives on {self.w.device}, input on {dev}: move the module first")
        d = _lib.StftDesc()
        d.nfft = nfft
        d.frame_width = self.win_length
        d.hop = self.frame_hop
        d.center_pad = self.center_pad()
        d.rescale = int(rescale)
        d.utt_preemph = float(utt_preemph)
        pre = float(self.pre_emphasis)
        d.frame_preemph = pre
        d.frame_one_minus = float(np.float32(1 - pre))
        if self.inverse:
            d.scale = nfft**-0.5 if self.normalized else 1.0 / nfft
        else:
            d.scale = nfft**-0.5 if self.normalized else 1.0
        win = self.kernel_window()
        d.window = win.data_ptr()
        d.twiddles = _lib.fft_tables(nfft, self.inverse, dev).data_ptr()
        d._keep = (win,)
        return d

    def num_frames(self, wav_len: th.Tensor) -> th.Tensor:
        """Frame count (integer exact; mutates `wav_len` in place when center=True — utils.py:653-662)."""
        assert th.sum(wav_len <= self.win_length) == 0
        if self.center:
            wav_len += self.win_length
        return th.div(wav_len - self.win_length, self.frame_hop, rounding_mode="trunc") + 1

    def extra_repr(self) -> str:
        s = (f"num_bins={self.num_bins}, win_length={self.win_length}, stride={self.frame_hop}, "
             f"window={self.window}, center={self.center}, mode={self.mode}")
        if not self.onesided:
            s += f", onesided={self.onesided}"
        if self.pre_emphasis > 0:
            s += f", pre_emphasis={self.pre_emphasis}"
        if self.normalized:
            s += f", normalized={self.normalized}"
        return s


class STFT(STFTBase):
    """Short-time Fourier transform layer: N x (C) x S -> N x (C) x F x T x 2 (utils.py:678-717)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, inverse=False, **kwargs)

    def forward(self, wav: th.Tensor, return_polar: bool = False, eps: float = EPSILON) -> th.Tensor:
        return stft_forward(self, wav, return_polar=return_polar, eps=eps)


def stft_forward(layer: STFTBase, wav: th.Tensor, return_polar: bool = False, eps: float = EPSILON,
                 rescale: bool = False, utt_preemph: float = 0.0) -> th.Tensor:
    """F2 launch.  `rescale` / `utt_preemph` fold a preceding RescaleTransform / PreEmphasisTransform."""
    if wav.dim() not in (2, 3):
        raise RuntimeError(f"STFT expect 2D/3D tensor, but got {wav.dim():d}D")
    dev = _lib.require_cuda(wav, "the STFT input")
    x = wav.detach()
    if x.dtype != th.float32:
        x = x.float()
    S = x.shape[-1]
    rows = x.numel() // S
    x = x.reshape(rows, S)
    if x.stride(-1) != 1:
        x = x.contiguous()
    desc = layer.stft_desc(dev, rescale=rescale, utt_preemph=utt_preemph)
    lib = _lib.load()
    T = lib.aps_b200_num_frames(S, desc.frame_width, desc.hop, desc.center_pad)
    if T < 1:
        raise RuntimeError(f"STFT: {S} samples are too few for one frame of {desc.frame_width}")
    out = th.empty((rows, layer.num_bins, T, 2), dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        _lib.check(lib.aps_b200_stft_fwd(x.data_ptr(), rows, S, x.stride(0), desc, int(return_polar), float(eps),
                                         out.data_ptr(), _lib.stream_ptr(dev)))
    if layer.mode == "torch" and wav.dim() == 3:
        # reference quirk (utils.py:405-407): `N` is re-bound to N*C before the view, so the 3-D
        # input comes back as [N*C, 1, F, T, 2] in torch mode
        return out.view(rows, 1, layer.num_bins, T, 2)
    return out.view(*wav.shape[:-1], layer.num_bins, T, 2)


class iSTFT(STFTBase):
    """Inverse STFT layer: (N) x F x T x 2 -> N x S (utils.py:720-758)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, inverse=True, **kwargs)

    def forward(self, transform: th.Tensor, return_polar: bool = False, eps: float = EPSILON) -> th.Tensor:
        if transform.dim() == 3:
            transform = transform[None]
        if transform.dim() != 4:
            raise RuntimeError(f"Expect 4D tensor, but got {transform.dim()}D")
        dev = _lib.require_cuda(transform, "the iSTFT input")
        x = transform.detach()
        if x.dtype != th.float32:
            x = x.float()
        x = x.contiguous()
        N, F, T, two = x.shape
        if F != self.num_bins or two != 2:
            raise RuntimeError(f"iSTFT expects N x {self.num_bins} x T x 2, got {tuple(x.shape)}")
        desc = self.stft_desc(dev)
        lib = _lib.load()
        S = lib.aps_b200_istft_num_samples(T, desc.frame_width, desc.hop, desc.center_pad)
        if S < 1:
            raise RuntimeError("iSTFT: not enough frames")
        out = th.empty((N, S), dtype=th.float32, device=dev)
        # th.istft (mode="torch") divides by the plain window envelope, the dense path by (envelope + eps)
        keps = 0.0 if self.mode == "torch" else float(eps)
        with th.cuda.device(dev):
            _lib.check(lib.aps_b200_istft_fwd(x.data_ptr(), N, T, desc, int(return_polar), keps, out.data_ptr(),
                                              _lib.stream_ptr(dev)))
        return out


def forward_stft(wav: th.Tensor, frame_len: int, frame_hop: int, return_polar: bool = False,
                 window: str = "sqrthann", round_pow_of_two: bool = True, pre_emphasis: float = 0,
                 normalized: bool = False, onesided: bool = True, center: bool = False, mode: str = "librosa",
                 eps: float = EPSILON) -> th.Tensor:
    """Functional STFT (utils.py:472-528)."""
    layer = STFT(frame_len, frame_hop, window=window, round_pow_of_two=round_pow_of_two,
                 pre_emphasis=pre_emphasis, normalized=normalized, onesided=onesided, center=center,
                 mode=mode).to(wav.device)
    return layer(wav, return_polar=return_polar, eps=eps)


def inverse_stft(transform: th.Tensor, frame_len: int, frame_hop: int, return_polar: bool = False,
                 window: str = "sqrthann", round_pow_of_two: bool = True, normalized: bool = False,
                 onesided: bool = True, center: bool = False, mode: str = "librosa",
                 eps: float = EPSILON) -> th.Tensor:
    """Functional iSTFT (utils.py:531-591)."""
    layer = iSTFT(frame_len, frame_hop, window=window, round_pow_of_two=round_pow_of_two,
                  normalized=normalized, onesided=onesided, center=center, mode=mode).to(transform.device)
    return layer(transform, return_polar=return_polar, eps=eps)
