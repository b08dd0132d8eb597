"""Oracle (CPU, fp32) for rows a14–a18 of SURVEY.md §8: mask post-processing, masked spatial
covariance, channel-attention reference vector, Souden MVDR weights via the real 2C x 2C inverse,
and beamforming.  TEST INFRASTRUCTURE — see oracle/__init__.py.

Complex tensors are (real, imag) pairs of real tensors exactly like the reference's
`aps.cplx.ComplexTensor` (aps/cplx.py:18-33); the arithmetic follows the reference operation by
operation (four real matmuls per complex matmul, cplx.py:242-252; inverse through the real block
matrix [[Re, -Im], [Im, Re]], cplx.py:268-278) so that its rounding behaviour is comparable.
"""
from typing import Optional, Tuple

import torch as th
import torch.nn.functional as F

from .transform import F32_EPS

Cplx = Tuple[th.Tensor, th.Tensor]


def cmatmul(a: Cplx, b: Cplx) -> Cplx:
    """Ref: aps/cplx.py:242-252 (_lmatmul)."""
    return (th.matmul(a[0], b[0]) - th.matmul(a[1], b[1]), th.matmul(a[1], b[0]) + th.matmul(a[0], b[1]))


def cinverse(a: Cplx) -> Cplx:
    """Ref: aps/cplx.py:268-278 (_inverse): invert [[Re, -Im], [Im, Re]], read back the top block row."""
    top = th.cat([a[0], -1.0 * a[1]], -1)
    bot = th.cat([a[1], a[0]], -1)
    inv = th.cat([top, bot], -2).inverse()
    C = a[0].shape[-1]
    return inv[..., :C, :C], -inv[..., :C, C:]


def padding_mask(lens: th.Tensor) -> th.Tensor:
    """True where t >= len.  Ref: aps/asr/base/attention.py:18-36."""
    return th.arange(int(lens.max()), device=lens.device)[None, :] >= lens[:, None]


def process_mask(mask: Optional[th.Tensor], x_len: Optional[th.Tensor], mask_norm: bool = True) -> Optional[th.Tensor]:
    """N x T x F -> N x F x T: zero the padded frames, divide by the per-(utt, bin) max over time.
    Ref: aps/asr/filter/mvdr.py:103-116."""
    if mask is None:
        return None
    if x_len is not None:
        mask = th.masked_fill(mask, padding_mask(x_len)[..., None], 0)
    if mask_norm:
        mask = mask / (th.norm(mask, float("inf"), dim=1, keepdim=True) + F32_EPS)
    return mask.transpose(1, 2)


def estimate_covar(mask: th.Tensor, spec: Cplx) -> Cplx:
    """mask N x F x T, spec N x C x F x T -> N x F x C x C.  Ref: mvdr.py:42-61."""
    xr, xi = spec[0].transpose(1, 2), spec[1].transpose(1, 2)        # N x F x C x T
    m = mask.unsqueeze(-2)
    num = cmatmul((xr * m, xi * m), (xr.transpose(-1, -2), -1.0 * xi.transpose(-1, -2)))
    den = th.clamp(m.sum(-1, keepdims=True), min=F32_EPS)
    return num[0] / den, num[1] / den


def channel_attention(Rs: Cplx, proj_w, proj_b, gvec_w, gvec_b) -> th.Tensor:
    """Reference-channel softmax u [N, C].  Ref: mvdr.py:148-174."""
    C = Rs[0].shape[-1]
    eye = th.eye(C, dtype=th.bool)
    r = Rs[0].masked_fill(eye, 0).sum(-1) / (C - 1)
    i = Rs[1].masked_fill(eye, 0).sum(-1) / (C - 1)
    a = (r**2 + i**2).sqrt().transpose(1, 2)                         # N x C x F
    g = F.linear(th.tanh(F.linear(a, proj_w, proj_b)), gvec_w, gvec_b)
    return F.softmax(g.squeeze(-1), -1)


def mvdr_weight(Rs: Cplx, Rn: Cplx, u: th.Tensor, eps: float = 1e-5) -> Cplx:
    """w = (Rn^-1 Rs) u / (tr(Rn^-1 Rs) + eps), N x F x C.  Ref: mvdr.py:75-101 (+ cplx.py:221-226 for
    the complex / complex division)."""
    C = Rn[0].shape[-1]
    Rn = (Rn[0] + th.eye(C) * eps, Rn[1])
    M = cmatmul(cinverse(Rn), Rs)
    eye = th.eye(C, dtype=th.bool).expand(*M[0].shape)
    tr = (M[0].masked_select(eye).view(*M[0].shape[:-1]).sum(-1) + eps,
          M[1].masked_select(eye).view(*M[1].shape[:-1]).sum(-1))
    nr = (M[0] * u[:, None, None, :]).sum(-1)
    ni = (M[1] * u[:, None, None, :]).sum(-1)
    tr_r, tr_i = tr[0][..., None], tr[1][..., None]
    scale = tr_r**2 + tr_i**2
    return (nr * tr_r + ni * tr_i) / scale, (ni * tr_r - nr * tr_i) / scale


def beamform(w: Cplx, spec: Cplx) -> Cplx:
    """w N x C x F, spec N x C x F x T -> N x F x T: sum_c conj(w) x.  Ref: mvdr.py:29-39."""
    wr, wi = w[0][..., None], -1.0 * w[1][..., None]
    return (wr * spec[0] - wi * spec[1]).sum(1), (wi * spec[0] + wr * spec[1]).sum(1)


def mvdr_forward(mask_s: th.Tensor, spec: Cplx, params: dict, mask_n: Optional[th.Tensor] = None,
                 x_len: Optional[th.Tensor] = None, mask_norm: bool = True, eps: float = 1e-5) -> Cplx:
    """MvdrBeamformer.forward: returns N x T x F.  Ref: mvdr.py:118-145.
    `params`: ref.proj.weight / ref.proj.bias / ref.gvec.weight / ref.gvec.bias."""
    ms = process_mask(mask_s, x_len, mask_norm)
    mn = process_mask(mask_n, x_len, mask_norm)
    Rs = estimate_covar(ms, spec)
    Rn = estimate_covar(1 - ms if mn is None else mn, spec)
    u = channel_attention(Rs, params["ref.proj.weight"], params["ref.proj.bias"], params["ref.gvec.weight"],
                          params["ref.gvec.bias"])
    w = mvdr_weight(Rs, Rn, u, eps)
    y = beamform((w[0].transpose(1, 2), w[1].transpose(1, 2)), spec)
    return y[0].transpose(1, 2), y[1].transpose(1, 2)
