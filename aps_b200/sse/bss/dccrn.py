"""DCCRN (`sse@dccrn`, complex variant) with the constructor, methods and `state_dict` layout of
/root/reference/aps/sse/bss/dccrn.py:139-323 (+ aps/sse/enh/dcunet.py:24-274), executed by the sm_100a
kernels through the C ABI.

Each complex (transposed) convolution — four real convolutions, two subtractions/additions and a
concatenation in the reference (dcunet.py:41-45) — is ONE real implicit-GEMM convolution on stacked
[real | imag] channels with the block weight [[Wr, -Wi], [Wi, Wr]], the per-part eval BatchNorm folded
into weight and bias and LeakyReLU in the epilogue.  Activations are channels-last [N, F, T, 2C]; the
packed STFT [N, F, T, 2] is already that layout for the first layer.  STFT / iSTFT are the F2 / F3
kernels, the complex ratio mask + mask application one small kernel.

The two-layer complex LSTM bottleneck (dccrn.py:20-110) keeps its parameters in `torch.nn.LSTM` modules
(reference `state_dict` layout) but runs through `ops.lstm`: one tensor-core GEMM per layer for the input
projections of all frames and a fused recurrence + cell-update launch per frame (csrc/lstm.cu).
Inference (`eval()`) only; `cplx=True`, `share_decoder=True`, non-causal convolutions.
"""
import os
from typing import List, Optional, Tuple, Union

import numpy as np
import torch as th
import torch.nn as nn

from ... import _lib, ops

EPSILON = float(np.finfo(np.float32).eps)


def parse_1dstr(s: str) -> List[int]:
    return list(map(int, s.split(",")))


def parse_2dstr(s: str) -> List[List[int]]:
    return [parse_1dstr(t) for t in s.split(";")]


# ------------------------------------------------------------------------------------ parameter containers
class ComplexConv2d(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.real, self.imag = nn.Conv2d(*a, **k), nn.Conv2d(*a, **k)


class ComplexConvTranspose2d(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.real, self.imag = nn.ConvTranspose2d(*a, **k), nn.ConvTranspose2d(*a, **k)


class ComplexBatchNorm2d(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.real_bn, self.imag_bn = nn.BatchNorm2d(*a, **k), nn.BatchNorm2d(*a, **k)


class EncoderBlock(nn.Module):
    """dcunet.py:103-143"""

    def __init__(self, cin, cout, kernel, stride, padding):
        super().__init__()
        self.kernel, self.stride = tuple(kernel), tuple(stride)
        self.padding = (padding, (kernel[-1] - 1) // 2)
        self.block = nn.Sequential(ComplexConv2d(cin, cout, self.kernel, stride=self.stride, padding=self.padding),
                                   ComplexBatchNorm2d(cout), nn.LeakyReLU())


class DecoderBlock(nn.Module):
    """dcunet.py:146-185"""

    def __init__(self, cin, cout, kernel, stride, padding, output_padding, last_layer):
        super().__init__()
        self.kernel, self.stride = tuple(kernel), tuple(stride)
        tpad = (kernel[-1] - 1) // 2
        self.padding, self.output_padding = (padding, kernel[1] - 1 - tpad), (output_padding, 0)
        mods = [ComplexConvTranspose2d(cin, cout, self.kernel, stride=self.stride, padding=self.padding,
                                       output_padding=self.output_padding)]
        if not last_layer:
            mods += [ComplexBatchNorm2d(cout), nn.LeakyReLU()]
        self.block = nn.Sequential(*mods)
        self.last = last_layer


class Encoder(nn.Module):
    def __init__(self, K, S, C, P):
        super().__init__()
        self.layers = nn.ModuleList([EncoderBlock(C[i], C[i + 1], k, S[i], P[i]) for i, k in enumerate(K)])


class Decoder(nn.Module):
    def __init__(self, K, S, C, P, O, connection):
        super().__init__()
        if connection not in ("cat", "sum"):
            raise ValueError(f"Unknown connection mode: {connection}")
        self.layers = nn.ModuleList([
            DecoderBlock(C[i] * 2 if connection == "cat" and i != 0 else C[i], C[i + 1], k, S[i], P[i], O[i],
                         last_layer=(i == len(K) - 1)) for i, k in enumerate(K)
        ])


class LSTMP(nn.Module):
    def __init__(self, in_features, hidden_size, num_layers=2, dropout=0, bidirectional=False):
        super().__init__()
        self.lstm = nn.LSTM(in_features, hidden_size, dropout=dropout, num_layers=num_layers,
                            bidirectional=bidirectional, batch_first=True)
        self.proj = nn.Linear(hidden_size * 2 if bidirectional else hidden_size, in_features, bias=False)

    def project(self, out: th.Tensor) -> th.Tensor:
        N, T, H = out.shape
        return ops.linear(out.reshape(N * T, H), self.proj.weight.detach()).view(N, T, -1)

    def splits(self):
        if not hasattr(self, "_splits"):
            self._splits = ops.SplitCache()
        return self._splits

    def forward(self, x: th.Tensor) -> th.Tensor:            # N x T x D
        return lstmp_pair([self], [x])[0]


def lstmp_pair(mods, xs):
    """Several LSTMP modules of one shape on same-shaped inputs: input projections of all frames on the tensor-core
    engine, then one fused recurrence + cell-update launch per frame for all of them together (csrc/lstm.cu), exact
    fp32.  There is no library path: hidden sizes that are not a multiple of 4 run zero padded (ops.lstm_multi)."""
    outs = ops.lstm_multi(xs, [m.lstm for m in mods], [m.splits() for m in mods])
    return [m.project(o) for m, o in zip(mods, outs)]


class ComplexLSTMP(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.real, self.imag = LSTMP(*a, **k), LSTMP(*a, **k)


class LSTMWrapper(nn.Module):
    def __init__(self, in_features, num_layers=2, dropout=0, hidden_size=512, bidirectional=False):
        super().__init__()
        self.lstm = ComplexLSTMP(in_features, hidden_size, dropout=dropout, num_layers=num_layers,
                                 bidirectional=bidirectional)


def _bn_scale_shift(bn):
    s = bn.weight.detach() / th.sqrt(bn.running_var + bn.eps)
    return s, bn.bias.detach() - bn.running_mean * s


def _pack_complex(real, imag, cbn, transposed: bool):
    """-> (weight [2Co, KH, KW, 2Ci], bias [2Co]) of the equivalent real convolution on [re | im] channels."""
    wr, wi = real.weight.detach(), imag.weight.detach()
    if transposed:                                           # [Ci, Co, KH, KW] -> [Co, Ci, KH, KW]
        wr, wi = wr.transpose(0, 1), wi.transpose(0, 1)
    top = th.cat([wr, -wi], 1)                               # yr = Wr*xr - Wi*xi
    bot = th.cat([wi, wr], 1)                                # yi = Wi*xr + Wr*xi
    w = th.cat([top, bot], 0)
    b = th.cat([real.bias.detach() - imag.bias.detach(), imag.bias.detach() + real.bias.detach()])
    if cbn is not None:
        sr, tr = _bn_scale_shift(cbn.real_bn)
        si, ti = _bn_scale_shift(cbn.imag_bn)
        s, t = th.cat([sr, si]), th.cat([tr, ti])
        w = w * s.view(-1, 1, 1, 1)
        b = b * s + t
    return w.permute(0, 2, 3, 1).contiguous(), b.contiguous()


_cat_complex = ops.cat_complex


class DCCRN(nn.Module):
    """Deep Complex Convolutional-RNN network (arguments as in dccrn.py:150-168)."""

    def __init__(self, cplx: bool = True, K: str = "3,3;3,3;3,3;3,3;3,3;3,3;3,3",
                 S: str = "2,1;2,1;2,1;2,1;2,1;2,1;2,1", P: str = "1,1,1,1,1,1,1", O: str = "0,0,0,0,0,0,0",
                 C: str = "16,32,64,64,128,128,256", num_spks: int = 2, connection: str = "sum", rnn_hidden: int = 512,
                 rnn_layers: int = 2, rnn_resize: int = 1536, rnn_dropout: float = 0, rnn_bidir: bool = False,
                 causal_conv: bool = False, share_decoder: bool = True, enh_transform: Optional[nn.Module] = None,
                 non_linear: str = "tanh", training_mode: str = "time") -> None:
        super().__init__()
        assert enh_transform is not None
        assert training_mode in ("freq", "time")
        if not cplx or causal_conv or not share_decoder:
            raise RuntimeError("aps_b200.DCCRN implements cplx=True, causal_conv=False, share_decoder=True")
        if non_linear not in ("none", "relu", "tanh", "sigmoid"):
            raise ValueError(f"Unsupported nonlinear: {non_linear}")
        self.enh_transform, self.training_mode = enh_transform, training_mode
        self.cplx, self.non_linear = cplx, non_linear
        self.forward_stft = enh_transform.ctx(name="forward_stft")
        self.inverse_stft = enh_transform.ctx(name="inverse_stft")
        K, S, C, P, O = parse_2dstr(K), parse_2dstr(S), parse_1dstr(C), parse_1dstr(P), parse_1dstr(O)
        self.encoder = Encoder(K, S, [1] + C, P)
        C = list(C)
        if connection == "cat":
            C[-1] *= 2
        self.decoder = nn.ModuleList([Decoder(K[::-1], S[::-1], C[::-1] + [num_spks], P[::-1], O[::-1], connection)])
        self.rnn = LSTMWrapper(rnn_resize // 2, dropout=rnn_dropout, num_layers=rnn_layers, hidden_size=rnn_hidden,
                               bidirectional=rnn_bidir)
        self.num_spks, self.connection, self.share_decoder = num_spks, connection, share_decoder
        self._packs = None
        self._splits = ops.SplitCache()
        self._guard = ops.PackGuard(self)
        self.register_load_state_dict_post_hook(lambda m, k: m._reset())

    def _reset(self):
        self._packs = None
        self._splits.clear()
        self._guard.reset()

    def refresh_packs(self):
        """Drop every derived copy of the weights (call after writing parameters through `.data`)."""
        self._reset()

    def _apply(self, fn, *a, **k):
        self._packs = None
        if hasattr(self, "_splits"):
            self._splits.clear()
            self._guard.reset()
        return super()._apply(fn, *a, **k)

    def _build_packs(self):
        enc = [_pack_complex(b.block[0].real, b.block[0].imag, b.block[1], False) for b in self.encoder.layers]
        dec = [_pack_complex(b.block[0].real, b.block[0].imag, None if b.last else b.block[1], True)
               for b in self.decoder[0].layers]
        return {"enc": enc, "dec": dec}

    # ---- mask estimation: packed STFT [N, F, T, 2] -> mask channels [N, F, T, 2*spks] ----------------------
    def _mask_nhwc(self, packed: th.Tensor) -> th.Tensor:
        if self.training:
            raise RuntimeError("aps_b200.DCCRN implements the inference forward only: call .eval()")
        if self._packs is None or self._guard.stale():       # also catches in-place parameter updates (ops.PackGuard)
            self._reset()
            self._packs = self._build_packs()
            self._guard.mark()
        pk = self._packs
        x = packed
        skips = []
        L = len(self.encoder.layers)
        for i, (blk, (w, b)) in enumerate(zip(self.encoder.layers, pk["enc"])):
            x = ops.conv2d_nhwc(x, w, b, stride=blk.stride, padding=blk.padding, act="leaky_relu", leaky=0.01,
                                cache=self._splits)
            if i + 1 != L:
                skips.append(x)
        # ---- complex LSTM bottleneck (csrc/lstm.cu): features ordered (channel, frequency) like dccrn.py:41-50 ------
        N, Fq, T, C2 = x.shape
        Cc = C2 // 2
        hr = x[..., :Cc].permute(0, 2, 3, 1).reshape(N, T, Cc * Fq)
        hi = x[..., Cc:].permute(0, 2, 3, 1).reshape(N, T, Cc * Fq)
        # complex LSTM (dccrn.py:97-110): out_r = R(hr) - I(hi), out_i = R(hi) + I(hr).  The two applications of each
        # real LSTM run as ONE call on the batch-concatenated inputs (rows of a batch are independent in an LSTM), and
        # the two LSTMs advance together, one launch per frame for both
        R, I = self.rnn.lstm.real, self.rnn.lstm.imag
        r_all, i_all = lstmp_pair([R, I], [th.cat([hr, hi], 0), th.cat([hi, hr], 0)])
        out_r = r_all[:N] - i_all[:N]
        out_i = r_all[N:] + i_all[N:]
        back = lambda t: t.view(N, T, Cc, Fq).permute(0, 3, 1, 2)          # -> N x F x T x Cc
        out = th.cat([back(out_r), back(out_i)], -1)
        # "cat" skip connections are not materialised: the transposed-conv gather reads both tensors in place
        cat = self.connection != "sum"
        x, other = (out, x) if cat else (x + out, None)
        skips = skips[::-1]
        for i, (blk, (w, b)) in enumerate(zip(self.decoder[0].layers, pk["dec"])):
            if i:
                x, other = (x, skips[i - 1]) if cat else (x + skips[i - 1], None)
            x = ops.conv_transpose2d_nhwc(x.contiguous(), w, b, stride=blk.stride, padding=blk.padding,
                                          output_padding=blk.output_padding,
                                          act="none" if blk.last else "leaky_relu", leaky=0.01, cache=self._splits,
                                          skip=other.contiguous() if other is not None else None)
        return x

    def _infer(self, mix: th.Tensor, mode: str):
        packed = self.forward_stft(mix, return_polar=False)                  # N x F x T x 2
        masks = self._mask_nhwc(packed)
        if masks.shape[:3] != packed.shape[:3]:
            raise RuntimeError(f"decoder output {tuple(masks.shape)} does not match the STFT {tuple(packed.shape)}; "
                               "check K/S/P/O (the reference's default P/O do not round-trip either, Q18)")
        outs = []
        for s in range(self.num_spks):
            m = ops.cmask(masks, s, self.num_spks + s, packed if mode == "time" else None, self.non_linear, EPSILON,
                          mode == "time")
            outs.append(self.inverse_stft(m, return_polar=False) if mode == "time" else m)
        return outs[0] if self.num_spks == 1 else outs

    def infer(self, mix: th.Tensor, mode: str = "time"):
        if mix.dim() != 1:
            raise RuntimeError(f"Expects 1D tensor (inference), got {mix.dim()} instead")
        with th.no_grad():
            sep = self._infer(mix[None, :], mode=mode)
            return sep[0] if self.num_spks == 1 else [s[0] for s in sep]

    def forward(self, s: th.Tensor):
        if s.dim() != 2:
            raise RuntimeError(f"Expects 2D tensor (training), got {s.dim()} instead")
        return self._infer(s, mode=self.training_mode)
