"""Oracle (CPU, fp32) for row a24 of SURVEY.md §8: DCCRN (`sse@dccrn`) forward in eval mode, evaluated
functionally from a reference `state_dict`.  TEST INFRASTRUCTURE — see oracle/__init__.py.
Follows /root/reference/aps/sse/bss/dccrn.py:16-294 and aps/sse/enh/dcunet.py:24-274 (complex
convolution = four real convolutions on the two halves of the frequency axis, per-part BatchNorm,
LeakyReLU, U-Net with `cat`/`sum` skips, complex LSTM bottleneck, complex ratio mask)."""
from typing import Dict, List

import torch as th
import torch.nn.functional as F

from .transform import F32_EPS, dft_kernel, istft_dense, stft_dense, window


def _cconv(sd, pre, x, stride, padding, transposed=False, output_padding=(0, 0)):
    """dcunet.py:24-66"""
    xr, xi = th.chunk(x, 2, -2)
    if transposed:
        op = lambda t, part: F.conv_transpose2d(t, sd[pre + part + ".weight"], sd[pre + part + ".bias"], stride=stride,
                                                padding=padding, output_padding=output_padding)
    else:
        op = lambda t, part: F.conv2d(t, sd[pre + part + ".weight"], sd[pre + part + ".bias"], stride=stride,
                                      padding=padding)
    return th.cat([op(xr, "real") - op(xi, "imag"), op(xr, "imag") + op(xi, "real")], -2)


def _cbn(sd, pre, x):
    """dcunet.py:72-87"""
    xr, xi = th.chunk(x, 2, -2)
    bn = lambda t, p: F.batch_norm(t, sd[pre + p + ".running_mean"], sd[pre + p + ".running_var"], sd[pre + p + ".weight"],
                                   sd[pre + p + ".bias"], False, 0.0, 1e-5)
    return th.cat([bn(xr, "real_bn"), bn(xi, "imag_bn")], -2)


def _lstmp(sd, pre, x, layers):
    """dccrn.py:16-50 (unidirectional): x N x T x C x F"""
    N, T, C, _ = x.shape
    h = x.reshape(N, T, -1)
    for l in range(layers):
        w_ih, w_hh = sd[f"{pre}lstm.weight_ih_l{l}"], sd[f"{pre}lstm.weight_hh_l{l}"]
        b_ih, b_hh = sd[f"{pre}lstm.bias_ih_l{l}"], sd[f"{pre}lstm.bias_hh_l{l}"]
        H = w_hh.shape[1]
        hs, cs = th.zeros(N, H), th.zeros(N, H)
        outs = []
        gi = F.linear(h, w_ih, b_ih)
        for t in range(T):
            g = gi[:, t] + F.linear(hs, w_hh, b_hh)
            i, f, gg, o = th.chunk(g, 4, -1)
            cs = th.sigmoid(f) * cs + th.sigmoid(i) * th.tanh(gg)
            hs = th.sigmoid(o) * th.tanh(cs)
            outs.append(hs)
        h = th.stack(outs, 1)
    return F.linear(h, sd[pre + "proj.weight"]).view(N, T, C, -1)


def tf_mask(sd: Dict[str, th.Tensor], real: th.Tensor, imag: th.Tensor, K, S, P, O, connection: str,
            rnn_layers: int = 2) -> th.Tensor:
    """dccrn.py:260-294 (cplx=True, share_decoder=True): N x F x T (x2) -> masks N x spks x 2F x T"""
    x = th.cat([real, imag], -2)[:, None]
    L = len(K)
    enc_h = []
    for i in range(L):
        pre = f"encoder.layers.{i}.block."
        x = _cconv(sd, pre + "0.", x, tuple(S[i]), (P[i], (K[i][1] - 1) // 2))
        x = F.leaky_relu(_cbn(sd, pre + "1.", x))
        if i + 1 != L:
            enc_h.append(x)
    h = x.permute(0, 3, 1, 2)                                         # N x T x C x 2F
    hr, hi = th.chunk(h, 2, -1)
    R = lambda t: _lstmp(sd, "rnn.lstm.real.", t, rnn_layers)
    I = lambda t: _lstmp(sd, "rnn.lstm.imag.", t, rnn_layers)
    out = th.cat([R(hr) - I(hi), R(hi) + I(hr)], -1).permute(0, 2, 3, 1)
    x = x + out if connection == "sum" else th.cat([out, x], 1)
    enc_h = enc_h[::-1]
    Kd, Sd, Pd, Od = K[::-1], S[::-1], P[::-1], O[::-1]
    for i in range(L):
        pre = f"decoder.0.layers.{i}.block."
        if i:
            x = x + enc_h[i - 1] if connection == "sum" else th.cat([x, enc_h[i - 1]], 1)
        tpad = (Kd[i][1] - 1) // 2
        x = _cconv(sd, pre + "0.", x, tuple(Sd[i]), (Pd[i], Kd[i][1] - 1 - tpad), True, (Od[i], 0))
        if i != L - 1:
            x = F.leaky_relu(_cbn(sd, pre + "1.", x))
    return x


def separate(m: th.Tensor, sr: th.Tensor, si: th.Tensor, non_linear: str):
    """Complex ratio mask application (dccrn.py:217-232, mode="time" before the iSTFT)."""
    mr, mi = th.chunk(m, 2, -2)
    m_abs = (mr**2 + mi**2 + F32_EPS)**0.5
    m_mag = {"sigmoid": th.sigmoid, "tanh": th.tanh, "relu": th.relu, "none": lambda t: t}[non_linear](m_abs)
    mr, mi = m_mag * mr / m_abs, m_mag * mi / m_abs
    return th.stack([sr * mr - si * mi, sr * mi + si * mr], -1)


def forward(sd, mix: th.Tensor, K, S, P, O, connection="cat", non_linear="sigmoid", num_spks=1, frame_len=512,
            frame_hop=256, window_name="sqrthann", center=True) -> List[th.Tensor]:
    """DCCRN.forward in training_mode="time" (dccrn.py:244-258): N x S -> [N x S] per speaker."""
    w0 = window(window_name, frame_len)
    Kf, wf = dft_kernel(frame_len, w0)
    Ki, wi = dft_kernel(frame_len, w0, inverse=True)
    packed = stft_dense(mix, Kf, wf, frame_hop, center=center)
    sr, si = packed[..., 0], packed[..., 1]
    masks = tf_mask(sd, sr, si, K, S, P, O, connection)
    return [istft_dense(separate(masks[:, s], sr, si, non_linear), Ki, wi, frame_hop, center=center)
            for s in range(num_spks)]
