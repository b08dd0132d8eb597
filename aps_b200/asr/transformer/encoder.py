"""Transformer / Conformer encoder with the constructor, `forward(inp_pad, inp_len)` contract and
`state_dict` layout of /root/reference/aps/asr/transformer/encoder.py:18-106, executed by the sm_100a
kernels (csrc/gemm.cu, csrc/encoder.cu) through the C ABI.

The sub-modules below are PARAMETER CONTAINERS that reproduce the reference's module tree name for
name (`proj.conv.enc_layers.{i}.conv|norm.norm`, `proj.conv.outp`, `pose.embed|div_term`,
`encoder.layers.{l}.self_attn.in_proj_weight|…`, `…feedforward{1,2}.{0,3}`, `…convolution.{0,2,3,5}`,
`…norm_*`, `encoder.norm`, `outp` — aps/asr/transformer/impl.py, proj.py, pose.py,
aps/asr/base/encoder.py:368-441, component.py:251-307) so checkpoints load with strict=True.  The
forward pass never calls them: `TransformerEncoder.forward` runs a fused inference schedule on
batch-major token rows [N*T', D]:

  conv2d front   : NHWC implicit-GEMM convolutions with eval-BatchNorm folded into weight/bias + ReLU
  FFN            : GEMM(+bias+activation) -> GEMM(+bias) -> LayerNorm(alpha*y + x)
  attention      : GEMM (packed QKV) -> mhsa kernel (abs / rel / xl, masks) -> GEMM(+bias + residual)
  conv module    : LayerNorm -> GEMM with interleaved GLU epilogue -> depthwise conv (+BN folded, Swish)
                   -> GEMM(+bias + residual)

Inference only (`eval()`): train-mode BatchNorm statistics / dropout / autograd are outside the
forward hot path this package covers.
"""
import math
import os
from typing import Dict, List, Optional, Tuple

import torch as th
import torch.nn as nn

from ... import _lib, ops

MIN_F32 = th.finfo(th.float32).min


class Swish(nn.Module):
    def forward(self, x):
        return x * th.sigmoid(x)


def _activation(name: str) -> nn.Module:
    if name == "relu":
        return nn.ReLU()
    if name == "gelu":
        return nn.GELU()
    if name == "swish":
        return Swish()
    raise RuntimeError(f"activation should be relu/gelu, not {name}")


def _relative_uv(shape) -> nn.Parameter:
    p = nn.Parameter(th.empty(*shape))
    nn.init.xavier_uniform_(p)
    return p


# ------------------------------------------------------------------------------------ containers: front
class Normalize2d(nn.Module):
    def __init__(self, name: str, features: int):
        super().__init__()
        name = name.upper()
        if name not in ("BN", "IN"):
            raise ValueError(f"Unknown type of Normalize2d: {name}")
        self.norm = nn.BatchNorm2d(features) if name == "BN" else nn.InstanceNorm2d(features)


class Normalize1d(nn.Module):
    def __init__(self, name: str, features: int):
        super().__init__()
        name = name.upper()
        if name not in ("BN", "LN"):
            raise ValueError(f"Unknown type of Normalize1d: {name}")
        self.norm = nn.BatchNorm1d(features) if name == "BN" else nn.GroupNorm(1, features)


class Conv2d(nn.Module):
    """Conv2d -> Norm -> ReLU block (component.py:251-307)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=2, dilation=1, norm="BN",
                 for_streaming=False):
        super().__init__()
        two = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        kernel_size, dilation = two(kernel_size), two(dilation)
        padding = tuple((d * (k - 1)) // 2 for d, k in zip(dilation, kernel_size))
        if for_streaming:
            padding = (0, padding[-1])
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation)
        self.norm = Normalize2d(norm, out_channels)
        self.kernel_size, self.padding, self.dilation, self.stride = kernel_size, padding, dilation, two(stride)

    def compute_outp_dim(self, dim: th.Tensor, axis: int) -> th.Tensor:
        """The reference's own length rule, `dim + 2p - d*k` (component.py:290-297, Q17) — kept verbatim."""
        return th.div(dim + 2 * self.padding[axis] - self.dilation[axis] * self.kernel_size[axis],
                      self.stride[axis], rounding_mode="trunc") + 1


class Conv2dEncoder(nn.Module):
    """Stack of Conv2d blocks + output projection (aps/asr/base/encoder.py:368-441)."""

    def __init__(self, inp_features, out_features, channel=32, in_channels=1, norm="BN", num_layers=3, kernel=3,
                 stride=2, for_streaming=False):
        super().__init__()

        def per_layer(v):
            if isinstance(v, int):
                return [(v, v)] * num_layers
            return [(p, p) for p in v] if isinstance(v[0], int) else list(v)

        self.kernel, self.stride = per_layer(kernel), per_layer(stride)
        channel = [channel] * num_layers if isinstance(channel, int) else channel
        self.enc_layers = nn.ModuleList([
            Conv2d(in_channels if i == 0 else channel[i - 1], channel[i], kernel_size=self.kernel[i], norm=norm,
                   stride=self.stride[i], for_streaming=for_streaming) for i in range(num_layers)
        ])
        freq = th.IntTensor([inp_features])
        for c in self.enc_layers:
            freq = c.compute_outp_dim(freq, 1)
        fxc = freq.item() * channel[-1]
        self.inp_features = inp_features
        if out_features > 0:
            self.out_features = out_features
            self.outp = nn.Linear(fxc, out_features)
        else:
            self.out_features = fxc
            self.outp = None


class Conv2dProj(nn.Module):
    """proj.py:104-140"""

    def __init__(self, input_size, embed_dim, norm="BN", kernel=3, stride=2, num_layers=2, in_channels=1,
                 conv_channels=256, for_streaming=False):
        super().__init__()
        assert num_layers in (2, 3, 4)
        self.conv = Conv2dEncoder(input_size, embed_dim, channel=conv_channels, in_channels=in_channels,
                                  num_layers=num_layers, norm=norm, kernel=kernel, stride=stride,
                                  for_streaming=for_streaming)


class LinearProj(nn.Module):
    """proj.py:31-57"""

    def __init__(self, input_size, embed_dim, dropout=0.0, norm="LN"):
        super().__init__()
        self.proj = nn.Linear(input_size, embed_dim)
        self.norm = Normalize1d(norm, embed_dim)
        self.drop = nn.Dropout(p=dropout)


# ------------------------------------------------------------------------------------ containers: pose
class SinPosEncoding(nn.Module):
    """pose.py:28-62 ("xl") / :93-122 ("abs")"""

    def __init__(self, embed_dim, dropout=0.0, scaled=False):
        super().__init__()
        div = th.exp(-math.log(10000.0) * th.arange(0, embed_dim, 2.0) / embed_dim)
        self.div_term = nn.Parameter(div, requires_grad=False)
        self.dropout = nn.Dropout(p=dropout)
        self.factor = embed_dim**0.5 if scaled else 1

    def encode(self, position: th.Tensor) -> th.Tensor:
        seq = position[:, None] * self.div_term
        return th.stack([th.sin(seq), th.cos(seq)], -1).view(position.shape[0], -1)


class RelPosEncoding(nn.Module):
    """pose.py:65-90"""

    def __init__(self, embed_dim, dropout=0.0, lradius=128, rradius=128):
        super().__init__()
        self.embed = nn.Embedding(lradius + rradius + 1, embed_dim)
        self.dropout = nn.Dropout(p=dropout)
        self.lradius, self.rradius = lradius, rradius

    def encode(self, position: th.Tensor) -> th.Tensor:
        position = th.clamp(position, max=self.rradius, min=-self.lradius)
        return self.embed.weight.detach()[position + self.lradius]


# ------------------------------------------------------------------------------------ containers: layers
class MultiheadAttention(nn.Module):
    """Parameters of ApsMultiheadAttention / Rel… / Xl… (impl.py:22-50, :299-322)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0, xl=False, rel_u=None, rel_v=None):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(th.empty(3 * embed_dim, embed_dim))
        nn.init.xavier_uniform_(self.in_proj_weight)
        self.in_proj_bias = nn.Parameter(th.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.dropout = nn.Dropout(p=dropout)
        if xl:
            shape = (num_heads, self.head_dim)
            self.rel_u = _relative_uv(shape) if rel_u is None else rel_u
            self.rel_v = _relative_uv(shape) if rel_v is None else rel_v
            self.rel_proj = nn.Linear(embed_dim, embed_dim, bias=False)


def _ffn(att_dim, ffn_dim, dropout, activation):
    return nn.Sequential(nn.Linear(att_dim, ffn_dim), _activation(activation), nn.Dropout(dropout),
                         nn.Linear(ffn_dim, att_dim), nn.Dropout(dropout))


class TransformerLayer(nn.Module):
    """impl.py:377-429"""

    def __init__(self, att_dim, self_attn, feedforward_dim=2048, dropout=0.1, activation="relu", pre_norm=False):
        super().__init__()
        self.self_attn = self_attn
        self.feedforward = _ffn(att_dim, feedforward_dim, dropout, activation)
        self.norm1 = nn.LayerNorm(att_dim)
        self.norm2 = nn.LayerNorm(att_dim)
        self.dropout = nn.Dropout(dropout)
        self.pre_norm, self.activation = pre_norm, activation


class ConformerLayer(nn.Module):
    """impl.py:432-497"""

    def __init__(self, att_dim, self_attn, feedforward_dim=2048, dropout=0.1, kernel_size=15, macaron=True,
                 pre_norm=True, casual_conv1d=False, activation="swish"):
        super().__init__()
        assert kernel_size % 2 == 1
        self.self_attn = self_attn
        if macaron:
            self.norm_ffn1 = nn.LayerNorm(att_dim)
            self.macaron_factor = 0.5
            self.feedforward1 = _ffn(att_dim, feedforward_dim, dropout, activation)
        else:
            self.macaron_factor = 1
            self.norm_ffn1 = None
            self.feedforward1 = None
        self.convolution = nn.Sequential(
            nn.Conv1d(att_dim, att_dim * 2, 1), nn.GLU(dim=-2),
            nn.Conv1d(att_dim, att_dim, kernel_size, groups=att_dim,
                      padding=0 if casual_conv1d else (kernel_size - 1) // 2), nn.BatchNorm1d(att_dim),
            _activation(activation), nn.Conv1d(att_dim, att_dim, 1), nn.Dropout(p=dropout))
        self.norm_ffn2 = nn.LayerNorm(att_dim)
        self.feedforward2 = _ffn(att_dim, feedforward_dim, dropout, activation)
        self.norm_attn = nn.LayerNorm(att_dim)
        self.norm_conv = nn.LayerNorm(att_dim)
        self.dropout = nn.Dropout(dropout)
        self.padding = kernel_size - 1 if casual_conv1d else 0
        self.kernel_size = kernel_size
        self.pre_norm, self.activation = pre_norm, activation


class LayerStack(nn.Module):
    """ApsTransformerEncoder (impl.py:718-756): `layers` + optional final `norm`."""

    def __init__(self, layers: List[nn.Module], norm: Optional[nn.Module]):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.num_layers = len(layers)
        self.norm = norm


def _build_layers(arch: str, pose: str, num_layers: int, kw: Dict) -> LayerStack:
    """get_xfmr_encoder (impl.py:759-787)."""
    if arch not in ("xfmr", "cfmr") or pose not in ("abs", "rel", "xl"):
        raise ValueError(f"Unknown type of the encoders: {arch}_{pose}")
    kw = dict(kw)
    att_dim, nhead = kw.pop("att_dim"), kw.pop("nhead")
    att_dropout, ffn_dropout = kw.pop("att_dropout", 0.1), kw.pop("ffn_dropout", 0.1)
    tie = kw.pop("tie", False) if pose == "xl" else False
    kw.pop("rel_u", None), kw.pop("rel_v", None)
    # Reference quirk (impl.py:768): the final LayerNorm exists only when "pre_norm" is GIVEN and true — a conformer built
    # without the key has pre-norm layers (their own default, impl.py:445) but no final norm.  Found by tests/test_dropin.py.
    final_norm = nn.LayerNorm(att_dim) if kw.get("pre_norm", False) else None
    shared = (_relative_uv((nhead, att_dim // nhead)), _relative_uv((nhead, att_dim // nhead))) if tie else None
    layers = []
    for _ in range(num_layers):
        u, v = shared if shared else (None, None)
        attn = MultiheadAttention(att_dim, nhead, dropout=att_dropout, xl=(pose == "xl"), rel_u=u, rel_v=v)
        if arch == "xfmr":
            layers.append(TransformerLayer(att_dim, attn, dropout=ffn_dropout, **kw))
        else:
            layers.append(ConformerLayer(att_dim, attn, dropout=ffn_dropout, **kw))
    return LayerStack(layers, final_norm)


def _context_mask(T: int, chunk: int, lctx: int, rctx: int, device) -> th.Tensor:
    """Additive (-inf / 0) T x T context mask (aps/asr/transformer/utils.py:61-98)."""
    lctx = T if lctx < 0 else lctx
    rctx = T if rctx < 0 else rctx
    idx = th.arange(T, device=device)
    fl = th.div(idx, chunk, rounding_mode="floor")
    right = (fl + rctx + 1) * chunk
    left = th.clamp_min((fl - lctx) * chunk, 0)
    cols = idx[None, :].expand(T, T)
    bad = (cols >= right[:, None]) | (cols < left[:, None])
    return th.zeros(T, T, device=device).masked_fill(bad, float("-inf"))


# ------------------------------------------------------------------------------------ the encoder
class TransformerEncoder(nn.Module):
    """Transformer based encoders: arch {xfmr|cfmr}, pose {abs|rel|xl}, proj {conv2d|linear|none}."""

    def __init__(self, arch: str, input_size: int, output_proj: int = -1, num_layers: int = 6, lctx: int = -1,
                 rctx: int = -1, chunk_size: int = 1, proj: str = "conv2d", proj_kwargs: Dict = {},
                 pose: str = "abs", pose_kwargs: Dict = {}, arch_kwargs: Dict = {}):
        super().__init__()
        att_dim = arch_kwargs["att_dim"]
        if proj == "none":
            self.proj = None
        elif proj == "conv2d":
            self.proj = Conv2dProj(input_size, att_dim, **proj_kwargs)
        elif proj == "linear":
            self.proj = LinearProj(input_size, att_dim, **proj_kwargs)
        else:
            raise ValueError(f"Unsupported projection layer: {proj} (aps_b200 provides conv2d | linear | none)")
        if pose == "rel":
            self.pose = RelPosEncoding(att_dim // arch_kwargs["nhead"], **pose_kwargs)
        elif pose in ("abs", "xl"):
            self.pose = SinPosEncoding(att_dim, **pose_kwargs)
        else:
            raise ValueError(f"Unsupported pose layer: {pose} (aps_b200 provides abs | rel | xl)")
        self.pose_type, self.arch = pose, arch
        self.encoder = _build_layers(arch, pose, num_layers, arch_kwargs)
        self.lctx, self.rctx, self.chunk_size = lctx, rctx, chunk_size
        self.outp = nn.Linear(att_dim, output_proj) if output_proj > 0 else None
        self.att_dim, self.nhead = att_dim, arch_kwargs["nhead"]
        self._packs = None
        self._splits = ops.SplitCache()
        self._graphs = {}
        self._guard = ops.PackGuard(self)
        self._fast = False
        self._ragged_lens = None
        self.use_graphs = os.environ.get("APS_B200_GRAPHS", "1") != "0"
        self.register_load_state_dict_post_hook(lambda m, k: m._drop_packs())

    # ---- repacked weights (BatchNorm folding, layout changes) ----------------------------------------------
    # Rebuilt after load_state_dict / .to() and whenever a parameter or buffer was updated in place since they were
    # built (ops.PackGuard: EMA / optimizer steps, a load_state_dict on a sub-module) — the captured CUDA graphs and the
    # TF32 splits they read go with them, so a replay can never see stale weights.
    def _drop_packs(self):
        self._packs = None
        self._splits.clear()
        self._graphs.clear()
        self._guard.reset()

    def refresh_packs(self):
        """Drop every derived copy of the weights (call after writing parameters through `.data`)."""
        self._drop_packs()

    def _apply(self, fn, *a, **k):
        self._packs = None
        if hasattr(self, "_splits"):
            self._splits.clear()
            self._graphs.clear()
            self._guard.reset()
        return super()._apply(fn, *a, **k)

    @staticmethod
    def _fold_bn(w: th.Tensor, b: Optional[th.Tensor], bn) -> Tuple[th.Tensor, th.Tensor]:
        scale = bn.weight.detach() / th.sqrt(bn.running_var + bn.eps)
        shift = bn.bias.detach() - bn.running_mean * scale
        b0 = b.detach() if b is not None else th.zeros_like(shift)
        return w.detach() * scale.view(-1, *([1] * (w.dim() - 1))), b0 * scale + shift

    def _build_packs(self, dev):
        pk = {"dev": dev, "layers": []}
        if isinstance(self.proj, LinearProj):
            nrm = self.proj.norm.norm
            if isinstance(nrm, nn.BatchNorm1d):
                w, b = self._fold_bn(self.proj.proj.weight, self.proj.proj.bias, nrm)
                pk["lin_w"], pk["lin_b"], pk["lin_ln"] = w.contiguous(), b.contiguous(), None
            else:   # "LN" = GroupNorm(1, D) over (D, T) of every utterance (component.py:95-96): a separate pass
                pk["lin_w"], pk["lin_b"] = self.proj.proj.weight.detach(), self.proj.proj.bias.detach()
                pk["lin_ln"] = (nrm.weight.detach(), nrm.bias.detach(), nrm.eps)
        if isinstance(self.proj, Conv2dProj):
            convs = []
            for blk in self.proj.conv.enc_layers:
                if not isinstance(blk.norm.norm, nn.BatchNorm2d):
                    raise RuntimeError("aps_b200: only norm='BN' is implemented for the conv2d projection")
                w, b = self._fold_bn(blk.conv.weight, blk.conv.bias, blk.norm.norm)
                convs.append((w.permute(0, 2, 3, 1).contiguous(), b.contiguous(), blk.stride, blk.padding))
            pk["convs"] = convs
            outp = self.proj.conv.outp
            if outp is not None:
                C = self.proj.conv.enc_layers[-1].conv.out_channels
                Fq = outp.in_features // C
                # reference flattens [C, F'] (channel-major); our activations are NHWC, i.e. [F', C]
                pk["front_w"] = outp.weight.detach().view(-1, C, Fq).permute(0, 2, 1).reshape(outp.out_features, -1).contiguous()
                pk["front_b"] = outp.bias.detach()
        for lay in self.encoder.layers:
            d = {}
            if isinstance(lay, ConformerLayer):
                c = lay.convolution
                D = c[0].in_channels
                w0 = c[0].weight.detach().view(2 * D, D)
                d["pw1_w"] = th.stack([w0[:D], w0[D:]], 1).reshape(2 * D, D).contiguous()      # interleave for GLU
                d["pw1_b"] = th.stack([c[0].bias.detach()[:D], c[0].bias.detach()[D:]], 1).reshape(-1).contiguous()
                wdw, bdw = self._fold_bn(c[2].weight, c[2].bias, c[3])
                d["dw_w"] = wdw.view(D, -1).t().contiguous()                                   # [K, D]
                d["dw_b"] = bdw.contiguous()
                d["pw2_w"] = c[5].weight.detach().view(D, D).contiguous()
                d["pw2_b"] = c[5].bias.detach()
            pk["layers"].append(d)
        return pk

    def _lin(self, x, w, b=None, **kw):
        return ops.linear(x, w, b, cache=self._splits, **kw)

    # ---- activations travel as PAIRS (x, x_lo) ----------------------------------------------------------------
    # x_lo is the TF32 "lo" companion of x (ops.linear): a tensor-core GEMM whose input has one loads both operand sides
    # by TMA and needs no gather / split warps.  Every kernel of the stack that feeds a GEMM writes the companion of its
    # output for free (GEMM / LayerNorm / attention / depthwise-conv epilogues); x_lo is None on the generic path
    # (SIMT engine, widths that are not multiples of 128), where everything below degrades to the plain kernels.
    def _lin2(self, xp, w, b=None, **kw):
        return ops.linear(xp[0], w, b, cache=self._splits, x_lo=xp[1], **kw)

    def _ln2(self, norm, x, residual=None, alpha=1.0, bias=None):
        """(y, y_lo) with y = LN(alpha * (sum of the split-K slices of x + bias) + residual); norm None: no LN."""
        if self._fast:
            g, b, eps = (norm.weight.detach(), norm.bias.detach(), norm.eps) if norm is not None else (None, None, 0.0)
            return ops.layernorm2(x, g, b, eps, bias=bias, residual=residual, alpha=alpha, normalize=norm is not None)
        assert x.dim() == 2 and bias is None and norm is not None
        return ops.layernorm(x, norm.weight.detach(), norm.bias.detach(), norm.eps, residual=residual, alpha=alpha), None

    def _ksplit(self, K: int, N: int, M: int) -> int:
        """K slices of the skinny GEMMs (d_ff -> d_model, front projection): 25 row tiles x 5 slices fill the machine at
        the BASELINE size.  A function of (K, N) only — never of M — so a row's result does not depend on the batch."""
        if not self._fast or M < 64 or N > 256 or K < 1024:
            return 1
        return 5 if K <= 2560 else 8

    def _second(self, hp, lin, residual, alpha, norm):
        """(y, y_lo): y = [LN](alpha * (h @ W.T + b) + residual) for the second Linear of an FFN."""
        w, b = lin.weight.detach(), lin.bias.detach()
        ks = self._ksplit(w.shape[1], w.shape[0], hp[0].shape[0])
        if ks > 1 and hp[1] is not None:
            parts = ops.linear(hp[0], w, None, cache=self._splits, x_lo=hp[1], ksplit=ks)
            return self._ln2(norm, parts, residual=residual, alpha=alpha, bias=b)
        if norm is None:
            return self._lin2(hp, w, b, alpha=alpha, residual=residual), None
        return self._ln2(norm, self._lin2(hp, w, b), residual=residual, alpha=alpha)

    def _ffn(self, seq, xp, act):
        """first Linear + activation of an FFN -> pair"""
        r = self._lin2(xp, seq[0].weight.detach(), seq[0].bias.detach(), act=act, want_lo=self._fast)
        return r if self._fast else (r, None)

    # ---- pieces ---------------------------------------------------------------------------------------------
    def _front_lens(self, lens: Optional[th.Tensor]) -> Optional[th.Tensor]:
        """Lengths after the projection front (integer exact, the reference's own rule)."""
        if lens is not None and isinstance(self.proj, Conv2dProj):
            for blk in self.proj.conv.enc_layers:
                lens = blk.compute_outp_dim(lens, 0)
        return lens

    def _front(self, x: th.Tensor, pk, in_lens: Optional[th.Tensor] = None):
        """-> (rows pair, N, T): token rows [N*T, D] after the projection front.  `in_lens` (device int64 [N], ragged-exact
        mode): frames beyond an utterance's length are zeroed before every convolution, i.e. each utterance sees the
        zero padding it would see alone (aps/asr/ctc.py:58-84 runs utterances one by one for exactly this reason)."""
        if self.proj is None:
            N, T, _ = x.shape
            return (ops.rows2d(x), None), N, T
        if isinstance(self.proj, LinearProj):
            N, T, Fi = x.shape
            if pk["lin_ln"] is None:
                y = self._lin(ops.rows2d(x), pk["lin_w"], pk["lin_b"], act="relu", want_lo=self._fast)
                return (y if self._fast else (y, None)), N, T
            g, b, eps = pk["lin_ln"]
            y = ops.utt_norm(self._lin(ops.rows2d(x), pk["lin_w"], pk["lin_b"]), N, T, g, b, eps, relu=True,
                             inplace=True)
            return (y, None), N, T
        x4 = x[:, None] if x.dim() == 3 else x                      # N x C x T x F
        nhwc = x4.permute(0, 2, 3, 1).contiguous()
        lo = None
        nconv = len(pk["convs"])
        lens = in_lens

        def zero_tail(t, ln):                                       # [N, T, F, C]: frames >= ln[n] <- 0
            keep = th.arange(t.shape[1], device=t.device)[None, :] < ln[:, None]
            return t * keep[:, :, None, None].to(t.dtype)

        if lens is not None:
            nhwc = zero_tail(nhwc, lens)
        for i, ((w, b, stride, padding), blk) in enumerate(zip(pk["convs"], self.proj.conv.enc_layers)):
            last = i + 1 == nconv and "front_w" in pk and self._fast
            nhwc = ops.conv2d_nhwc(nhwc, w, b, stride=stride, padding=padding, act="relu", cache=self._splits, want_lo=last)
            if last:
                nhwc, lo = nhwc
            if lens is not None:
                lens = blk.compute_outp_dim(lens, 0)
                if i + 1 < nconv:                                   # the projection after the last layer is per token
                    nhwc = zero_tail(nhwc, lens)
        N, T, Fq, C = nhwc.shape
        flat = nhwc.view(N * T, Fq * C)
        if "front_w" in pk:
            flo = lo.view(N * T, Fq * C) if lo is not None else None
            w, b = pk["front_w"], pk["front_b"]
            ks = self._ksplit(w.shape[1], w.shape[0], N * T)
            if ks > 1 and flo is not None:
                parts = ops.linear(flat, w, None, cache=self._splits, x_lo=flo, ksplit=ks)
                return self._ln2(None, parts, bias=b), N, T
            y = ops.linear(flat, w, b, cache=self._splits, x_lo=flo, want_lo=self._fast)
            return (y if self._fast else (y, None)), N, T
        # no output projection: restore the reference's channel-major feature order
        return (nhwc.permute(0, 1, 3, 2).reshape(N * T, C * Fq), None), N, T

    def _attention(self, a, xp, res, N, T, inj, kpm, amask):
        """res + SelfAttention(x) with the parameters of container `a` (plain tensor)."""
        qkv = self._lin2(xp, a.in_proj_weight.detach(), a.in_proj_bias.detach())
        lo = self._fast
        if self.pose_type == "rel":
            ctx = ops.mhsa(qkv, N, T, self.nhead, mode=1, pos=inj, kpm=kpm, kpm_fill=MIN_F32, attn_mask=amask, want_lo=lo)
        elif self.pose_type == "xl":
            pos = self._lin(inj, a.rel_proj.weight.detach())
            ctx = ops.mhsa(qkv, N, T, self.nhead, mode=2, pos=pos, rel_u=a.rel_u.detach().contiguous(),
                           rel_v=a.rel_v.detach().contiguous(), kpm=kpm, kpm_fill=MIN_F32, attn_mask=amask,
                           qpos_is_value=True, want_lo=lo)
        else:
            ctx = ops.mhsa(qkv, N, T, self.nhead, mode=0, kpm=kpm, kpm_fill=float("-inf"), attn_mask=amask, want_lo=lo)
        cp = ctx if lo else (ctx, None)
        return self._lin2(cp, a.out_proj.weight.detach(), a.out_proj.bias.detach(), residual=res)

    def _xfmr_layer(self, lay, pk, xp, N, T, inj, kpm, amask):
        act = lay.activation
        if lay.pre_norm:
            x = self._attention(lay.self_attn, self._ln2(lay.norm1, xp[0]), xp[0], N, T, inj, kpm, amask)
            hp = self._ffn(lay.feedforward, self._ln2(lay.norm2, x), act)
            return self._second(hp, lay.feedforward[3], x, 1.0, None)
        xp = self._ln2(lay.norm1, self._attention(lay.self_attn, xp, xp[0], N, T, inj, kpm, amask))
        hp = self._ffn(lay.feedforward, xp, act)
        return self._second(hp, lay.feedforward[3], xp[0], 1.0, lay.norm2)

    def _conv_module(self, lay, d, up, res, N, T):
        """conv(u) + res (plain tensor)"""
        g = self._lin2(up, d["pw1_w"], d["pw1_b"], act="glu")
        K = lay.kernel_size
        c = ops.dwconv1d(g, N, T, d["dw_w"], d["dw_b"], dilation=1, left_pad=(K - 1) if lay.padding else (K - 1) // 2,
                         act=lay.activation, want_lo=self._fast, lens=self._ragged_lens)
        return self._lin2(c if self._fast else (c, None), d["pw2_w"], d["pw2_b"], residual=res)

    def _cfmr_layer(self, lay, d, xp, N, T, inj, kpm, amask):
        act, mac = lay.activation, lay.macaron_factor
        if lay.pre_norm:
            x = xp[0]
            if lay.feedforward1 is not None:
                hp = self._ffn(lay.feedforward1, self._ln2(lay.norm_ffn1, x), act)
                x = self._second(hp, lay.feedforward1[3], x, mac, None)[0]
            x = self._attention(lay.self_attn, self._ln2(lay.norm_attn, x), x, N, T, inj, kpm, amask)
            x = self._conv_module(lay, d, self._ln2(lay.norm_conv, x), x, N, T)
            hp = self._ffn(lay.feedforward2, self._ln2(lay.norm_ffn2, x), act)
            return self._second(hp, lay.feedforward2[3], x, mac, None)
        if lay.feedforward1 is not None:
            hp = self._ffn(lay.feedforward1, xp, act)
            xp = self._second(hp, lay.feedforward1[3], xp[0], mac, lay.norm_ffn1)
        x = self._attention(lay.self_attn, xp, xp[0], N, T, inj, kpm, amask)
        x = self._conv_module(lay, d, self._ln2(lay.norm_attn, x), x, N, T)       # impl.py:536 reuses norm_attn
        xp = self._ln2(lay.norm_conv, x)
        hp = self._ffn(lay.feedforward2, xp, act)
        return self._second(hp, lay.feedforward2[3], xp[0], mac, lay.norm_ffn2)

    # ---- forward --------------------------------------------------------------------------------------------
    def _run(self, x: th.Tensor, lens_dev: Optional[th.Tensor], in_lens: Optional[th.Tensor] = None) -> th.Tensor:
        """Device-only part of the forward (safe to capture in a CUDA graph): x N x Ti x F -> N x To x D.
        `in_lens` (input frames per utterance, device): ragged-exact mode, see `forward_ragged`."""
        pk, dev = self._packs, x.device
        self._ragged_lens = lens_dev if in_lens is not None else None
        xp, N, T = self._front(x, pk, in_lens)
        D = xp[0].shape[-1]
        kpm = None
        if lens_dev is not None:
            kpm = (th.arange(T, device=dev)[None, :] >= lens_dev[:, None]).to(th.uint8).contiguous()
        inj = None
        if self.pose_type == "abs":
            x3 = xp[0].view(N, T, D) * self.pose.factor + self.pose.encode(th.arange(0, T, 1.0, device=dev))
            rows = x3.reshape(N * T, D).contiguous()
            xp = (rows, ops.lo_companion(rows) if self._fast else None)
        elif self.pose_type == "rel":
            inj = self.pose.encode(th.arange(-T + 1, T, device=dev)).contiguous()
        else:
            inj = self.pose.encode(th.arange(0, 2 * T - 1, 1.0, device=dev)).contiguous()
        amask = None
        if self.lctx != -1 or self.rctx != -1:
            amask = _context_mask(T, self.chunk_size, self.lctx, self.rctx, dev).contiguous()
        if not xp[0].is_contiguous():
            xp = (xp[0].contiguous(), None)
        if self._fast and xp[1] is None:
            xp = (xp[0], ops.lo_companion(xp[0]))
        for lay, d in zip(self.encoder.layers, pk["layers"]):
            if isinstance(lay, ConformerLayer):
                xp = self._cfmr_layer(lay, d, xp, N, T, inj, kpm, amask)
            else:
                xp = self._xfmr_layer(lay, d, xp, N, T, inj, kpm, amask)
        if self.encoder.norm is not None:
            xp = self._ln2(self.encoder.norm, xp[0])
        rows = xp[0]
        if self.outp is not None:
            rows = self._lin2(xp, self.outp.weight.detach(), self.outp.bias.detach())
        return rows.view(N, T, -1)

    def _run_graphed(self, x: th.Tensor, lens_dev: Optional[th.Tensor]) -> th.Tensor:
        """Replay a captured CUDA graph of `_run` once an input shape has been seen twice (the ~180 launches
        of a 12-layer forward otherwise cost more host time than device time at M = 3200 tokens)."""
        key = (tuple(x.shape), lens_dev is not None)
        state = self._graphs.get(key)
        if state is None:
            if len(self._graphs) >= 4:
                self._graphs.clear()
            self._graphs[key] = "seen"
            return self._run(x, lens_dev)
        if state == "seen":
            sx = x.clone()
            sl = lens_dev.clone() if lens_dev is not None else None
            graph = th.cuda.CUDAGraph()
            th.cuda.synchronize(x.device)
            try:
                with th.cuda.graph(graph):
                    out = self._run(sx, sl)
            except Exception:
                self.use_graphs = False
                self._graphs.clear()
                raise
            state = self._graphs[key] = (graph, sx, sl, out)
        graph, sx, sl, out = state
        sx.copy_(x)
        if sl is not None:
            sl.copy_(lens_dev)
        graph.replay()
        return out.clone()

    def _ensure_packs(self, dev):
        if self._packs is None or self._packs["dev"] != dev or self._guard.stale():
            if next(self.parameters()).device != dev:
                raise RuntimeError(f"encoder parameters on {next(self.parameters()).device}, input on {dev}")
            self._drop_packs()
            self._packs = self._build_packs(dev)
            self._guard.mark()
        # pair mode: tensor-core engine and a model width the fused LayerNorm / reduce kernel takes
        self._fast = (ops.GEMM_ENGINE == "tc" and ops.layernorm2_ok(self.att_dim)
                      and os.environ.get("APS_B200_ENC_PAIRS", "1") != "0")

    def ragged_exact_ok(self) -> bool:
        """Can a padded batch reproduce the one-utterance-at-a-time results?  Needs length-independent position terms
        ("xl" indexes its sinusoid table by T) and per-token normalisation in the front (LinearProj "LN" = GroupNorm over
        the utterance)."""
        if self.pose_type == "xl":
            return False
        if isinstance(self.proj, LinearProj) and not isinstance(self.proj.norm.norm, nn.BatchNorm1d):
            return False
        return True

    def forward_ragged(self, inp_pad: th.Tensor, inp_len: th.Tensor):
        """Padded batch N x Ti x F with lengths -> (N x To x D with rows beyond each length zeroed, lengths), where every
        utterance gets EXACTLY what `forward(inp_pad[n:n+1, :len[n]], None)` gives (up to the rounding of a different
        GEMM tile path): besides the key-padding mask, frames beyond an utterance's length are zeroed in front of every
        convolution (conv2d front, depthwise conv of the conformer layers) so that it sees the zero padding it would see
        alone.  The reference gets this by looping over the utterances (aps/asr/ctc.py:58-84); `forward` itself keeps the
        reference's batched semantics."""
        if self.training:
            raise RuntimeError("aps_b200.TransformerEncoder implements the inference forward only: call .eval()")
        if not self.ragged_exact_ok():
            raise RuntimeError("forward_ragged: this encoder configuration depends on the padded length (pose 'xl' or an "
                               "utterance-level norm in the projection); run the utterances one by one")
        dev = _lib.require_cuda(inp_pad, "encoder input")
        self._ensure_packs(dev)
        out_len = self._front_lens(inp_len)
        in_dev = inp_len.detach().to(device=dev, dtype=th.int64)
        lens_dev = out_len.detach().to(device=dev, dtype=th.int64)
        out = self._run(inp_pad.detach().float(), lens_dev, in_dev)
        self._ragged_lens = None
        keep = th.arange(out.shape[1], device=dev)[None, :] < lens_dev[:, None]
        return out * keep[:, :, None].to(out.dtype), out_len

    def forward(self, inp_pad: th.Tensor, inp_len: Optional[th.Tensor]):
        """inp_pad N x Ti x F (or N x C x Ti x F), inp_len N or None -> (N x To x D, lengths)."""
        if self.training:
            raise RuntimeError("aps_b200.TransformerEncoder implements the inference forward only: call .eval()")
        dev = _lib.require_cuda(inp_pad, "encoder input")
        self._ensure_packs(dev)
        x = inp_pad.detach().float()
        out_len = self._front_lens(inp_len)
        lens_dev = None
        if out_len is not None:
            lens_dev = out_len.detach().to(device=dev, dtype=th.int64)
        if self.use_graphs and not th.cuda.is_current_stream_capturing():
            out = self._run_graphed(x.contiguous(), lens_dev)
        else:
            out = self._run(x, lens_dev)
        if out_len is not None and int(out_len.max()) != out.shape[1]:
            raise RuntimeError(f"padding mask length {int(out_len.max())} does not match {out.shape[1]} encoder frames")
        return out, out_len
