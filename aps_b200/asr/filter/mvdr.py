"""Mask-based MVDR front-end with the surface of /root/reference/aps/asr/filter/mvdr.py
(`estimate_covar` :42, `beamform` :29, `ChannelAttention` :148, `MvdrBeamformer` :64), running on the
sm_100a kernels of csrc/mvdr.cu through the C ABI.

Parameters keep the reference's names (`ref.proj.{weight,bias}`, `ref.gvec.{weight,bias}`) so that
checkpoints load strictly.  Complex tensors go in and out as (real, imag) pairs of whatever class the
caller used (normally `aps.cplx.ComplexTensor`, built from the two halves of a packed STFT).
"""
from typing import Optional

import numpy as np
import torch as th
import torch.nn as nn

from ... import _lib
from ...cplx import is_complex_pair, like

EPSILON = float(np.finfo(np.float32).eps)   # aps/const.py:17


def _view(x, what: str):
    """(real, imag, dev, strides, shape) of an N x C x F x T complex pair, without copying when the two
    parts share strides (e.g. the halves of a packed STFT)."""
    if not is_complex_pair(x):
        raise RuntimeError(f"{what} must be a complex (real, imag) pair, got {type(x)}")
    re, im = x.real.detach(), x.imag.detach()
    dev = _lib.require_cuda(re, what)
    if re.dim() != 4 or re.shape != im.shape:
        raise RuntimeError(f"{what}: expect N x C x F x T, got {tuple(re.shape)} / {tuple(im.shape)}")
    if re.dtype != th.float32 or im.dtype != th.float32:
        re, im = re.float(), im.float()
    if re.stride() != im.stride():
        re, im = re.contiguous(), im.contiguous()
    C = re.shape[1]
    if not 2 <= C <= 6:
        raise RuntimeError(f"aps_b200 MVDR kernels support 2..6 channels, got {C}")
    return re, im, dev, _lib.i64_array(re.stride()), re.shape


def _mask_view(mask: th.Tensor, dev, order: str):
    """(tensor, strides {n, t, f}) of a real mask given as N x T x F ("ntf") or N x F x T ("nft")."""
    m = mask.detach()
    if m.dtype != th.float32:
        m = m.float()
    if m.device != dev:
        raise RuntimeError(f"mask on {m.device}, spectrogram on {dev}")
    sn, s1, s2 = m.stride()
    return m, _lib.i64_array((sn, s1, s2) if order == "ntf" else (sn, s2, s1))


def _covar(mask_s, mask_n, x, order: str, lens, normalise: bool, want_rn: bool):
    re, im, dev, xs, (N, C, F, T) = _view(x, "spectrogram")
    lib = _lib.load()
    st = _lib.stream_ptr(dev)
    ms, mss = _mask_view(mask_s, dev, order)
    exp = (N, T, F) if order == "ntf" else (N, F, T)
    if tuple(ms.shape) != exp:
        raise RuntimeError(f"mask shape {tuple(ms.shape)} does not match spectrogram {(N, C, F, T)}")
    mn = mns = None
    if mask_n is not None:
        mn, mns = _mask_view(mask_n, dev, order)
    lens_d = None
    if lens is not None:
        lens_d = lens.detach().to(device=dev, dtype=th.int64).contiguous()
    max_s = max_n = None
    with th.cuda.device(dev):
        if normalise:
            max_s = th.empty((N, F), dtype=th.float32, device=dev)
            _lib.check(lib.aps_b200_mask_colmax(ms.data_ptr(), mss[0], mss[1], mss[2], N, T, F, _lib.ptr(lens_d),
                                                max_s.data_ptr(), st))
            if mn is not None:
                max_n = th.empty((N, F), dtype=th.float32, device=dev)
                _lib.check(lib.aps_b200_mask_colmax(mn.data_ptr(), mns[0], mns[1], mns[2], N, T, F,
                                                    _lib.ptr(lens_d), max_n.data_ptr(), st))
        Rs = th.empty((N, F, C, C, 2), dtype=th.float32, device=dev)
        Rn = th.empty((N, F, C, C, 2), dtype=th.float32, device=dev) if want_rn else None
        _lib.check(lib.aps_b200_covar_fwd(re.data_ptr(), im.data_ptr(), xs, N, C, F, T, ms.data_ptr(), mss,
                                          _lib.ptr(max_s), _lib.ptr(mn), mns, _lib.ptr(max_n), _lib.ptr(lens_d),
                                          EPSILON, EPSILON, Rs.data_ptr(), _lib.ptr(Rn), st))
    return Rs, Rn


def estimate_covar(mask: th.Tensor, spectrogram):
    """mask N x F x T (real), spectrogram N x C x F x T (complex) -> covariance N x F x C x C (complex)
    (mvdr.py:42-61)."""
    Rs, _ = _covar(mask, None, spectrogram, "nft", None, False, False)
    return like(spectrogram, Rs[..., 0], Rs[..., 1])


def beamform(weight, spectrogram):
    """weight N x C x F, spectrogram N x C x F x T -> N x F x T: sum_c conj(w) x (mvdr.py:29-39)."""
    re, im, dev, xs, (N, C, F, T) = _view(spectrogram, "spectrogram")
    w = th.stack([weight.real, weight.imag], -1).detach().float().permute(0, 2, 1, 3).contiguous()  # N x F x C x 2
    if tuple(w.shape) != (N, F, C, 2):
        raise RuntimeError(f"weight shape {tuple(weight.real.shape)} does not match spectrogram {(N, C, F, T)}")
    yr = th.empty((N, F, T), dtype=th.float32, device=dev)
    yi = th.empty((N, F, T), dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_beamform_fwd(re.data_ptr(), im.data_ptr(), xs, N, C, F, T, w.data_ptr(),
                                                     yr.data_ptr(), yi.data_ptr(), _lib.stream_ptr(dev)))
    return like(spectrogram, yr, yi)


class ChannelAttention(nn.Module):
    """Reference-channel selection vector u (mvdr.py:148-174).  Holds the parameters; `forward` accepts
    the covariance as a complex pair N x F x C x C and returns softmax weights N x C."""

    def __init__(self, num_bins: int, att_dim: int) -> None:
        super().__init__()
        self.proj = nn.Linear(num_bins, att_dim)
        self.gvec = nn.Linear(att_dim, 1)

    def logits(self, Rs_packed: th.Tensor) -> th.Tensor:
        N, F, C = Rs_packed.shape[:3]
        dev = Rs_packed.device
        if self.proj.weight.device != dev:
            raise RuntimeError(f"ChannelAttention parameters on {self.proj.weight.device}, input on {dev}")
        out = th.empty((N, C), dtype=th.float32, device=dev)
        with th.cuda.device(dev):
            _lib.check(_lib.load().aps_b200_mvdr_ref_logits(
                Rs_packed.data_ptr(), N, F, C, self.proj.weight.detach().contiguous().data_ptr(),
                self.proj.bias.detach().data_ptr(), self.gvec.weight.detach().contiguous().data_ptr(),
                self.gvec.bias.detach().data_ptr(), self.proj.weight.shape[0], out.data_ptr(), _lib.stream_ptr(dev)))
        return out

    def forward(self, Rs) -> th.Tensor:
        packed = th.stack([Rs.real, Rs.imag], -1).detach().float().contiguous()
        _lib.require_cuda(packed, "covariance")
        return th.softmax(self.logits(packed), -1)


class MvdrBeamformer(nn.Module):
    """MVDR (minimum variance distortionless response) beamformer (mvdr.py:64-145)."""

    def __init__(self, num_bins, att_dim=512, mask_norm=True, eps=1e-5):
        super().__init__()
        self.ref = ChannelAttention(num_bins, att_dim)
        self.mask_norm = mask_norm
        self.eps = eps

    def forward(self, mask_s: th.Tensor, x, mask_n: Optional[th.Tensor] = None, x_len: Optional[th.Tensor] = None):
        """mask_s / mask_n: N x T x F real TF masks, x: N x C x F x T complex, x_len: N frame counts.
        Returns the enhanced complex spectrogram N x T x F."""
        Rs, Rn = _covar(mask_s, mask_n, x, "ntf", x_len, self.mask_norm, True)
        re, im, dev, xs, (N, C, F, T) = _view(x, "spectrogram")
        lib = _lib.load()
        st = _lib.stream_ptr(dev)
        logits = self.ref.logits(Rs)
        w = th.empty((N, F, C, 2), dtype=th.float32, device=dev)
        yr = th.empty((N, F, T), dtype=th.float32, device=dev)
        yi = th.empty((N, F, T), dtype=th.float32, device=dev)
        with th.cuda.device(dev):
            _lib.check(lib.aps_b200_mvdr_weights(Rs.data_ptr(), Rn.data_ptr(), logits.data_ptr(), N, F, C,
                                                 float(self.eps), w.data_ptr(), st))
            _lib.check(lib.aps_b200_beamform_fwd(re.data_ptr(), im.data_ptr(), xs, N, C, F, T, w.data_ptr(),
                                                 yr.data_ptr(), yi.data_ptr(), st))
        return like(x, yr.transpose(1, 2), yi.transpose(1, 2))
