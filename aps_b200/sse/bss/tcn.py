"""Frequency-domain Conv-TasNet (`sse@freq_tcn`) with the constructor, methods and `state_dict` layout of
/root/reference/aps/sse/bss/tcn.py:361-469, executed by the sm_100a kernels through the C ABI.

Module tree (names kept for strict checkpoint loading): `proj.1` (1x1 conv), `conv.repeat.{r}.{b}.`
{`conv1`, `norm1.0` (PReLU), `norm1.1` (norm), `dconv`, `norm2.0`, `norm2.1`, `conv2`} with optional
`.scale` scalars (ScaleLinear, tcn.py:91-109), `conv.skip_linear.{i}`, `mask.0` (PReLU), `mask.1`.

Fused inference schedule on batch-major token rows [N*T, C] — three launches per block:
  GEMM(1x1 conv * scale + bias, PReLU, BatchNorm affine)  ->  depthwise dilated conv (+bias, PReLU,
  BatchNorm affine)  ->  GEMM(1x1 conv * scale + bias, + residual)
With `norm="BN"` (the default of the frequency-domain model) the eval-mode affine folds into the epilogues as
above; "cLN" / "gLN" / "IN" need per-utterance statistics over time and add one `ops.utt_norm` (two small
launches) after each PReLU.
"""
import os
from typing import List, Optional, Union

import torch as th
import torch.nn as nn

from ... import _lib, ops


class ScaleLinear(nn.Conv1d):
    """1x1 Conv1d with an optional learnt output scale (parameter container)."""

    def __init__(self, in_features, out_features, bias=True, scale_param=1.0):
        super().__init__(in_features, out_features, 1, bias=bias)
        self.scale = nn.Parameter(th.tensor(scale_param)) if scale_param else 1

    def packed(self):
        s = self.scale.detach() if isinstance(self.scale, nn.Parameter) else 1.0
        w = (self.weight.detach()[..., 0] * s).contiguous()
        b = (self.bias.detach() * s).contiguous() if self.bias is not None else None
        return w, b


class GlobalChannelLayerNorm(nn.Module):
    """tcn.py:33-72 (parameter container: `beta`, `gamma` of shape [C, 1])."""

    def __init__(self, dim: int, eps: float = 1e-05, elementwise_affine: bool = True) -> None:
        super().__init__()
        self.eps, self.normalized_dim, self.elementwise_affine = eps, dim, elementwise_affine
        if elementwise_affine:
            self.beta = nn.Parameter(th.zeros(dim, 1))
            self.gamma = nn.Parameter(th.ones(dim, 1))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def extra_repr(self) -> str:
        return f"{self.normalized_dim}, eps={self.eps}, elementwise_affine={self.elementwise_affine}"


def _norm_layer(norm: str, channels: int) -> nn.Module:
    """tcn.py:75-88"""
    if norm not in ("cLN", "IN", "gLN", "BN"):
        raise RuntimeError(f"Unsupported normalize layer: {norm}")
    if norm == "cLN":
        return nn.GroupNorm(1, channels)
    if norm == "IN":
        return nn.GroupNorm(channels, channels)
    if norm == "BN":
        return nn.BatchNorm1d(channels)
    return GlobalChannelLayerNorm(channels)


def _norm_pack(m: nn.Module):
    """-> ("bn", (scale, shift)) folded eval affine, or ("utt", dict(gamma, beta, eps, per_channel))."""
    if isinstance(m, nn.BatchNorm1d):
        return "bn", _bn_affine(m)
    if isinstance(m, nn.GroupNorm):
        return "utt", dict(gamma=m.weight.detach() if m.affine else None, beta=m.bias.detach() if m.affine else None,
                           eps=m.eps, per_channel=m.num_groups != 1)
    g = m.gamma.detach().reshape(-1).contiguous() if m.elementwise_affine else None
    b = m.beta.detach().reshape(-1).contiguous() if m.elementwise_affine else None
    return "utt", dict(gamma=g, beta=b, eps=m.eps, per_channel=False)


class Conv1dBlock(nn.Module):
    """tcn.py:112-159 (parameter container)."""

    def __init__(self, in_channels=256, conv_channels=512, kernel_size=3, dilation=1, norm="cLN", scale_param=0,
                 causal=False):
        super().__init__()
        self.pad = dilation * (kernel_size - 1)
        self.cau, self.dilation, self.kernel_size = causal, dilation, kernel_size
        self.conv1 = ScaleLinear(in_channels, conv_channels, scale_param=scale_param)
        self.norm1 = nn.Sequential(nn.PReLU(), _norm_layer(norm, conv_channels))
        self.dconv = nn.Conv1d(conv_channels, conv_channels, kernel_size, groups=conv_channels,
                               padding=self.pad if causal else self.pad // 2, dilation=dilation)
        self.norm2 = nn.Sequential(nn.PReLU(), _norm_layer(norm, conv_channels))
        self.conv2 = ScaleLinear(conv_channels, in_channels, scale_param=scale_param)


class Conv1dRepeat(nn.Module):
    """tcn.py:162-226 (parameter container)."""

    def __init__(self, num_repeats, blocks_per_repeat, in_channels=128, conv_channels=128, kernel_size=3, norm="BN",
                 skip_residual=True, scaling_param=False, causal=False):
        super().__init__()
        self.repeat = nn.Sequential(*[
            nn.Sequential(*[
                Conv1dBlock(in_channels=in_channels, conv_channels=conv_channels, kernel_size=kernel_size, norm=norm,
                            causal=causal, dilation=2**n, scale_param=0 if scaling_param else 0.9**n)
                for n in range(blocks_per_repeat)
            ]) for _ in range(num_repeats)
        ])
        self.skip_residual = skip_residual
        if skip_residual:
            tot = num_repeats * (num_repeats - 1) // 2
            self.skip_linear = nn.ModuleList([ScaleLinear(in_channels, in_channels, scale_param=1.0)
                                              for _ in range(tot)])
        else:
            self.skip_linear = None


class _Transpose(nn.Module):
    def forward(self, x):
        return x.transpose(-1, -2)


def _bn_affine(bn: nn.BatchNorm1d):
    scale = bn.weight.detach() / th.sqrt(bn.running_var + bn.eps)
    return scale.contiguous(), (bn.bias.detach() - bn.running_mean * scale).contiguous()


def _pack_repeats(conv: Conv1dRepeat):
    """Kernel-ready parameters of a Conv1dRepeat stack (1x1 weights with the ScaleLinear scale folded in, depthwise
    weights tap-major, eval BatchNorm as an affine pair or the per-utterance norm description)."""
    pk = {"blocks": [], "skip": []}
    for rep in conv.repeat:
        for blk in rep:
            w1, b1 = blk.conv1.packed()
            w2, b2 = blk.conv2.packed()
            C = blk.dconv.weight.shape[0]
            pk["blocks"].append(dict(
                w1=w1, b1=b1, a1=blk.norm1[0].weight.detach(), n1=_norm_pack(blk.norm1[1]),
                wd=blk.dconv.weight.detach().view(C, -1).t().contiguous(), bd=blk.dconv.bias.detach(),
                a2=blk.norm2[0].weight.detach(), n2=_norm_pack(blk.norm2[1]), w2=w2, b2=b2,
                dil=blk.dilation, lpad=blk.pad if blk.cau else blk.pad // 2))
    if conv.skip_linear is not None:
        pk["skip"] = [lin.packed() for lin in conv.skip_linear]
    return pk


def _run_repeats(lin, conv: Conv1dRepeat, pk, x: th.Tensor, N: int, T: int) -> th.Tensor:
    """The repeat stack of tcn.py:162-226 on token rows [N*T, C]: three launches per block (see the module docstring)."""
    outs, skip, bi = [x], 0, 0
    nrep, nblk = len(conv.repeat), len(conv.repeat[0])
    # Pair schedule (as in the encoder): every activation travels with its TF32 lo companion, written by the kernel that
    # produces it, so both operand sides of the 1x1 convolutions arrive by TMA (no gather / split warps).  Needs
    # epilogue-only norms (eval BatchNorm or none) and no skip links; APS_B200_TCN_PAIRS=0 turns it off.
    pairs = (ops.GEMM_ENGINE == "tc" and not conv.skip_residual and x.shape[0] >= 128 and x.shape[1] % 4 == 0
             and all(d["n1"][0] != "utt" and d["n2"][0] != "utt" and d["w1"].shape[0] % 4 == 0 for d in pk["blocks"])
             and os.environ.get("APS_B200_TCN_PAIRS", "1") != "0")
    if pairs:
        x_lo = ops.lo_companion(x)
        for d in pk["blocks"]:
            (k1, n1), (k2, n2) = d["n1"], d["n2"]
            h, h_lo = lin(x, d["w1"], d["b1"], act="prelu", slope=d["a1"], post=n1 if k1 == "bn" else None, x_lo=x_lo,
                          want_lo=True)
            h, h_lo = ops.dwconv1d(h, N, T, d["wd"], d["bd"], dilation=d["dil"], left_pad=d["lpad"], act="prelu",
                                   slope=d["a2"], post=n2 if k2 == "bn" else None, want_lo=True)
            x, x_lo = lin(h, d["w2"], d["b2"], residual=x, x_lo=h_lo, want_lo=True)
        return x
    for r in range(nrep):
        if conv.skip_residual:
            for i in range(r):                       # in-place accumulation semantics of tcn.py:203-224
                w, b = pk["skip"][skip + i]
                x = lin(outs[i], w, b, residual=x)
            outs[r] = x
            skip += r
        for _ in range(nblk):
            d = pk["blocks"][bi]
            bi += 1
            (k1, n1), (k2, n2) = d["n1"], d["n2"]
            h = lin(x, d["w1"], d["b1"], act="prelu", slope=d["a1"], post=n1 if k1 == "bn" else None)
            if k1 == "utt":
                h = ops.utt_norm(h, N, T, inplace=True, **n1)
            h = ops.dwconv1d(h, N, T, d["wd"], d["bd"], dilation=d["dil"], left_pad=d["lpad"], act="prelu",
                             slope=d["a2"], post=n2 if k2 == "bn" else None)
            if k2 == "utt":
                h = ops.utt_norm(h, N, T, inplace=True, **n2)
            x = lin(h, d["w2"], d["b2"], residual=x)
        outs.append(x)
    return x


class FreqConvTasNet(nn.Module):
    """Frequency domain ConvTasNet (arguments as in tcn.py:366-381)."""

    def __init__(self, enh_transform: Optional[nn.Module] = None, in_features: int = 257, B: int = 6, K: int = 3,
                 N: int = 3, conv_channels: int = 512, proj_channels: int = 256, norm: str = "BN", num_spks: int = 2,
                 num_bins: int = 257, non_linear: str = "relu", causal: bool = False, scaling_param: bool = False,
                 skip_residual: bool = False, training_mode: str = "freq") -> None:
        super().__init__()
        assert enh_transform is not None
        assert training_mode in ("freq", "time")
        if non_linear not in ("relu", "sigmoid"):
            raise ValueError(f"Unsupported nonlinear: {non_linear}")
        self.enh_transform = enh_transform
        self.training_mode = training_mode
        self.proj = nn.Sequential(_Transpose(), nn.Conv1d(in_features, proj_channels, 1))
        self.conv = Conv1dRepeat(N, B, in_channels=proj_channels, conv_channels=conv_channels, kernel_size=K,
                                 causal=causal, scaling_param=scaling_param, skip_residual=skip_residual, norm=norm)
        self.mask = nn.Sequential(nn.PReLU(), nn.Conv1d(proj_channels, num_bins * num_spks, 1))
        self.non_linear = non_linear
        self.num_spks, self.num_bins = num_spks, num_bins
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

    def _lin(self, x, w, b=None, **kw):
        return ops.linear(x, w, b, cache=self._splits, **kw)

    def _build_packs(self):
        pk = _pack_repeats(self.conv)
        C = self.mask[1].in_channels
        pk["ident"] = th.ones(1, C, device=self.mask[1].weight.device)
        pk["proj_w"] = self.proj[1].weight.detach()[..., 0].contiguous()
        pk["mask_w"] = self.mask[1].weight.detach()[..., 0].contiguous()
        return pk

    def _mask_rows(self, feats: th.Tensor) -> th.Tensor:
        """feats N x T x F -> mask rows [N*T, num_bins*num_spks] (after the output non-linearity)."""
        if self.training:
            raise RuntimeError("aps_b200.FreqConvTasNet implements the inference forward only: call .eval()")
        dev = _lib.require_cuda(feats, "mask network input")
        if feats.dim() != 3:
            raise RuntimeError(f"expect N x T x F features, got {feats.dim()}D")
        if self._packs is None or self._guard.stale():       # also catches in-place parameter updates (ops.PackGuard)
            self._reset()
            self._packs = self._build_packs()
            self._guard.mark()
        pk = self._packs
        N, T, Fi = feats.shape
        rows = ops.rows2d(feats.detach().float())
        x = self._lin(rows, pk["proj_w"], self.proj[1].bias.detach())
        x = _run_repeats(self._lin, self.conv, pk, x, N, T)
        # mask head: PReLU then 1x1 conv then relu / sigmoid
        a = ops.dwconv1d(x, N, T, pk["ident"], None, act="prelu", slope=self.mask[0].weight.detach())
        return self._lin(a, pk["mask_w"], self.mask[1].bias.detach(), act=self.non_linear)

    def _tf_mask(self, feats: th.Tensor, num_spks: int) -> List[th.Tensor]:
        """[N x F x T, ...] (views of one N x T x (F*spks) buffer) — tcn.py:403-414."""
        N, T, _ = feats.shape
        m = self._mask_rows(feats).view(N, T, -1).transpose(1, 2)
        return list(th.chunk(m, self.num_spks, 1))

    def mask_predict(self, feats: th.Tensor) -> th.Tensor:
        """feats N x T x F -> masks N x F x T (or S x N x F x T) — tcn.py:458-469."""
        masks = th.stack(self._tf_mask(feats, self.num_spks))
        return masks[0] if self.num_spks == 1 else masks

    def _infer(self, mix: th.Tensor, mode: str):
        """tcn.py:416-431: STFT -> features -> masks (-> masked STFT -> iSTFT in "time" mode)."""
        stft, _ = self.enh_transform.encode(mix, None)
        masks = self._tf_mask(self.enh_transform(stft), self.num_spks)
        if mode == "time":
            ref = stft[:, 0] if stft.dim() == 5 else stft
            bss = self.enh_transform.decode([ref * m.unsqueeze(-1) for m in masks])   # aps/sse/base.py:23-49
        else:
            bss = masks
        return bss[0] if self.num_spks == 1 else bss

    def infer(self, mix: th.Tensor, mode: str = "time"):
        if mix.dim() not in (1, 2):
            raise RuntimeError(f"Expects 1/2D tensor (inference), got {mix.dim()} instead")
        with th.no_grad():
            ret = self._infer(mix[None, :], mode=mode)
            return ret[0] if self.num_spks == 1 else [r[0] for r in ret]

    def forward(self, mix: th.Tensor):
        if mix.dim() not in (2, 3):
            raise RuntimeError(f"Expects 2/3D tensor (training), got {mix.dim()} instead")
        return self._infer(mix, mode=self.training_mode)


class TimeConvTasNet(nn.Module):
    """Time-domain Conv-TasNet (`sse@time_tcn`, arguments as in tcn.py:241-255).

    Schedule on token rows [N*T, C] (T = (S - L) // (L/2) + 1 encoder frames):
      encoder Conv1d(1, N, L, stride L/2) + ReLU = a GEMM on the (overlapping) waveform frames, K = L  ->  cLN (`ops.utt_norm`)  ->  1x1 proj  ->  the shared repeat stack  ->
      PReLU, 1x1 mask conv with relu / sigmoid in the epilogue (softmax over speakers as one small device op)  ->
      w * m  ->  decoder ConvTranspose1d(N, 1, L, stride L/2) = a GEMM to L samples per frame + a two-term
      overlap-add (every output sample has exactly two contributing frames) + bias.
    Inference only; `mixture_consistency` other than "none" is marked "current not working" in the reference
    (tcn.py:305) and is refused here.
    """

    def __init__(self, L: int = 20, N: int = 256, X: int = 8, R: int = 4, B: int = 256, H: int = 512, P: int = 3,
                 norm: str = "BN", causal: bool = False, num_spks: int = 2, non_linear: str = "relu",
                 scaling_param: bool = False, skip_residual: bool = False, mixture_consistency: str = "none") -> None:
        super().__init__()
        assert mixture_consistency in ["none", "fix", "mag", "learn"]
        if non_linear not in ("relu", "sigmoid", "softmax"):
            raise ValueError(f"Unsupported nonlinear: {non_linear}")
        if mixture_consistency != "none":
            raise RuntimeError("aps_b200.TimeConvTasNet: mixture_consistency is not implemented "
                               "(the reference marks it as not working, tcn.py:305)")
        if L % 2:
            raise RuntimeError("aps_b200.TimeConvTasNet needs an even encoder length L")
        self.training_mode, self.enh_transform = "time", None
        self.non_linear_type = non_linear
        self.encoder = nn.Conv1d(1, N, L, stride=L // 2, padding=0)
        self.ln = _norm_layer("cLN", N)
        self.proj = nn.Conv1d(N, B, 1)
        self.conv = Conv1dRepeat(R, X, in_channels=B, conv_channels=H, kernel_size=P, norm=norm,
                                 skip_residual=skip_residual, scaling_param=scaling_param, causal=causal)
        self.mask = nn.Sequential(nn.PReLU(), nn.Conv1d(B, num_spks * N, 1))
        self.decoder = nn.ConvTranspose1d(N, 1, kernel_size=L, stride=L // 2, bias=True)
        self.num_spks, self.mixture_consistency, self.L = num_spks, mixture_consistency, L
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

    def _lin(self, x, w, b=None, **kw):
        return ops.linear(x, w, b, cache=self._splits, **kw)

    def _build_packs(self):
        pk = _pack_repeats(self.conv)
        dev = self.mask[1].weight.device
        pk["ident"] = th.ones(1, self.mask[1].in_channels, device=dev)
        pk["enc_w"] = self.encoder.weight.detach()[:, 0].contiguous()                    # [N, L]
        pk["proj_w"] = self.proj.weight.detach()[..., 0].contiguous()
        pk["mask_w"] = self.mask[1].weight.detach()[..., 0].contiguous()
        pk["dec_w"] = self.decoder.weight.detach()[:, 0].t().contiguous()                # [L, N]: frame samples x channels
        return pk

    def forward(self, mix: th.Tensor):
        """mix N x S -> [N x S', ...] (S' = (T - 1) * L/2 + L), tcn.py:326-358."""
        if mix.dim() != 2:
            raise RuntimeError(f"Expects 2D tensor (training), got {mix.dim()} instead")
        if self.training:
            raise RuntimeError("aps_b200.TimeConvTasNet implements the inference forward only: call .eval()")
        dev = _lib.require_cuda(mix, "mixture")
        if self._packs is None or self._guard.stale():       # also catches in-place parameter updates (ops.PackGuard)
            self._reset()
            self._packs = self._build_packs()
            self._guard.mark()
        pk = self._packs
        mix = mix.detach().float().contiguous()
        Nb, S = mix.shape
        L, hop = self.L, self.L // 2
        if S < L:
            raise RuntimeError(f"mixture of {S} samples is shorter than the encoder window {L}")
        T = (S - L) // hop + 1
        nf = self.encoder.out_channels
        # encoder: one GEMM per utterance batch over overlapping strided rows (row t = samples [t*hop, t*hop + L))
        frames = mix.unfold(-1, L, hop).reshape(Nb * T, L).contiguous()   # materialised once: 2x the waveform, K = L
        w = ops.linear(frames, pk["enc_w"], self.encoder.bias.detach(), act="relu")
        g = self.ln
        y = ops.utt_norm(w, Nb, T, g.weight.detach(), g.bias.detach(), g.eps)
        y = self._lin(y, pk["proj_w"], self.proj.bias.detach())
        y = _run_repeats(self._lin, self.conv, pk, y, Nb, T)
        a = ops.dwconv1d(y, Nb, T, pk["ident"], None, act="prelu", slope=self.mask[0].weight.detach())
        act = self.non_linear_type if self.non_linear_type != "softmax" else "none"
        e = self._lin(a, pk["mask_w"], self.mask[1].bias.detach(), act=act)              # [N*T, spks*nf]
        m = e.view(Nb * T, self.num_spks, nf)
        if self.non_linear_type == "softmax":
            m = th.softmax(m, 1)                                                          # over speakers (sse/base.py:50)
        bss = []
        To = (T - 1) * hop + L
        for s in range(self.num_spks):
            sw = (w * m[:, s]).contiguous()
            f = ops.linear(sw, pk["dec_w"]).view(Nb, T, L)                                # frame-wise output samples
            out = th.zeros((Nb, To), dtype=th.float32, device=dev)
            out[:, :T * hop].view(Nb, T, hop).add_(f[..., :hop])
            out[:, hop:(T + 1) * hop].view(Nb, T, hop).add_(f[..., hop:])
            bss.append(out + self.decoder.bias.detach())
        return bss[0] if self.num_spks == 1 else bss

    def infer(self, mix: th.Tensor, mode: str = "time"):
        if mix.dim() != 1:
            raise RuntimeError(f"Expects 1D tensor (inference), got {mix.dim()} instead")
        with th.no_grad():
            sep = self.forward(mix[None, ...])
            return sep[0] if self.num_spks == 1 else [s[0] for s in sep]
