"""Oracle (CPU, fp32) for rows a20–a23 of SURVEY.md §8: the transformer / conformer encoder forward
(`asr@…` `enc_type="xfmr"|"cfmr"`), evaluated functionally from a reference `state_dict`.
TEST INFRASTRUCTURE — see oracle/__init__.py.  Inference semantics only (dropout off, BatchNorm in
eval mode), which is what the reference's eval/decoding path runs.

The functions follow /root/reference/aps/asr/transformer/{encoder,impl,pose,proj,utils}.py and
/root/reference/aps/asr/base/{encoder,component}.py operation by operation, including the
reference's quirks: `Conv2d.compute_outp_dim` uses `dim + 2p - d*k` (component.py:290-297, Q17) and the
"xl" attention feeds VALUE where the query belongs (impl.py:369, Q14).
"""
import math
from typing import Dict, Optional, Tuple

import torch as th
import torch.nn.functional as F

MIN_F32 = th.finfo(th.float32).min


def padding_mask(lens: th.Tensor) -> th.Tensor:
    """aps/asr/base/attention.py:18-36"""
    return th.arange(int(lens.max()), device=lens.device)[None, :] >= lens[:, None]


def digit_shift(term: th.Tensor) -> th.Tensor:
    """L x N x H x (2L-1) -> L x N x H x L with out[l, ..., s] = term[l, ..., s - l + L - 1]
    (index form of aps/asr/transformer/utils.py:14-39, verified against it in the tests)."""
    L = term.shape[0]
    idx = th.arange(L)[None, :] - th.arange(L)[:, None] + L - 1          # [l, s]
    return th.gather(term, -1, idx[:, None, None, :].expand(L, term.shape[1], term.shape[2], L))


def sin_encoding(position: th.Tensor, div_term: th.Tensor) -> th.Tensor:
    """pose.py:41-50"""
    seq = position[:, None] * div_term
    return th.stack([th.sin(seq), th.cos(seq)], -1).view(position.shape[0], -1)


class EncoderOracle:
    """cfg: dict(arch, input_size, output_proj, num_layers, proj, proj_kwargs, pose, pose_kwargs,
    arch_kwargs, lctx, rctx, chunk_size) — the kwargs of aps/asr/transformer/encoder.py:23-35."""

    def __init__(self, cfg: Dict, sd: Dict[str, th.Tensor], prefix: str = ""):
        self.cfg = cfg
        self.sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        a = cfg["arch_kwargs"]
        self.D, self.H = a["att_dim"], a["nhead"]
        self.pre_norm = a.get("pre_norm", cfg["arch"] == "cfmr")
        self.act = a.get("activation", "swish" if cfg["arch"] == "cfmr" else "relu")
        self.pose = cfg.get("pose", "abs")

    def p(self, key: str) -> Optional[th.Tensor]:
        return self.sd.get(key)

    # ---- projection front (proj.py) ------------------------------------------------------------------
    def conv2d_proj(self, x: th.Tensor, lens: Optional[th.Tensor]):
        kw = self.cfg.get("proj_kwargs", {})
        nl = kw.get("num_layers", 2)
        x = x[:, None] if x.dim() == 3 else x
        for i in range(nl):
            pre = f"proj.conv.enc_layers.{i}."
            w = self.p(pre + "conv.weight")
            k = w.shape[-1]
            x = F.conv2d(x, w, self.p(pre + "conv.bias"), stride=2, padding=(k - 1) // 2)
            x = F.batch_norm(x, self.p(pre + "norm.norm.running_mean"), self.p(pre + "norm.norm.running_var"),
                             self.p(pre + "norm.norm.weight"), self.p(pre + "norm.norm.bias"), False, 0.0, 1e-5)
            x = F.relu(x)
            if lens is not None:          # component.py:290-297 (padding 1, dilation 1, kernel k, stride 2)
                lens = th.div(lens + 2 * ((k - 1) // 2) - k, 2, rounding_mode="trunc") + 1
        N, _, T, _ = x.shape
        x = x.transpose(1, 2).contiguous().view(N, T, -1)
        if self.p("proj.conv.outp.weight") is not None:
            x = F.linear(x, self.p("proj.conv.outp.weight"), self.p("proj.conv.outp.bias"))
        return x, lens

    def linear_proj(self, x, lens):
        x = F.linear(x, self.p("proj.proj.weight"), self.p("proj.proj.bias"))
        if self.p("proj.norm.norm.running_mean") is not None:      # BN over features (component.py:103-114)
            x = F.batch_norm(x.transpose(1, 2), self.p("proj.norm.norm.running_mean"),
                             self.p("proj.norm.norm.running_var"), self.p("proj.norm.norm.weight"),
                             self.p("proj.norm.norm.bias"), False, 0.0, 1e-5).transpose(1, 2)
        else:   # "LN" is nn.GroupNorm(1, C) on N x C x T: statistics over (C, T) of the utterance (component.py:96)
            x = F.group_norm(x.transpose(1, 2), 1, self.p("proj.norm.norm.weight"),
                             self.p("proj.norm.norm.bias")).transpose(1, 2)
        return F.relu(x), lens

    # ---- attention (impl.py) ---------------------------------------------------------------------------
    def attention(self, pre: str, x: th.Tensor, inj, kpm, amask) -> th.Tensor:
        L, N, E = x.shape
        H, dh = self.H, E // self.H
        qkv = F.linear(x, self.p(pre + "in_proj_weight"), self.p(pre + "in_proj_bias"))
        q, k, v = [t.view(L, N, H, dh) for t in th.chunk(qkv, 3, -1)]
        if self.pose == "rel":
            logit = th.einsum("lnhd,snhd->lnhs", q, k) + digit_shift(th.matmul(q, inj.transpose(0, 1)))
        elif self.pose == "xl":
            u, vv = self.p(pre + "rel_u"), self.p(pre + "rel_v")
            rel = F.linear(inj, self.p(pre + "rel_proj.weight")).view(-1, H, dh)
            x_ = v                                                    # impl.py:369 passes value as query
            logit = th.einsum("lnhd,snhd->lnhs", x_ + u, k) + digit_shift(th.einsum("lnhd,shd->lnhs", x_ + vv, rel))
        else:
            logit = th.einsum("lnhd,snhd->lnhs", q, k)
        logit = logit / dh**0.5
        fill = float("-inf") if self.pose == "abs" else MIN_F32      # torch MHA vs impl.py:105-107
        if kpm is not None:
            logit = logit.masked_fill(kpm[None, :, None, :], fill)
        if amask is not None:
            logit = logit + amask[:, None, None, :]
        ctx = th.einsum("lnhs,snhd->lnhd", th.softmax(logit, -1), v).contiguous().view(L, N, E)
        return F.linear(ctx, self.p(pre + "out_proj.weight"), self.p(pre + "out_proj.bias"))

    def _act(self, x):
        return {"relu": F.relu, "gelu": F.gelu, "swish": lambda t: t * th.sigmoid(t)}[self.act](x)

    def ffn(self, pre: str, x):
        h = self._act(F.linear(x, self.p(pre + "0.weight"), self.p(pre + "0.bias")))
        return F.linear(h, self.p(pre + "3.weight"), self.p(pre + "3.bias"))

    def ln(self, pre: str, x):
        return F.layer_norm(x, (x.shape[-1],), self.p(pre + "weight"), self.p(pre + "bias"))

    def xfmr_layer(self, i: int, x, inj, kpm, amask):
        """impl.py:402-429"""
        pre = f"encoder.layers.{i}."
        inp = self.ln(pre + "norm1.", x) if self.pre_norm else x
        x = x + self.attention(pre + "self_attn.", inp, inj, kpm, amask)
        if self.pre_norm:
            return x + self.ffn(pre + "feedforward.", self.ln(pre + "norm2.", x))
        x = self.ln(pre + "norm1.", x)
        return self.ln(pre + "norm2.", x + self.ffn(pre + "feedforward.", x))

    def conv_module(self, pre: str, x):
        """impl.py:483-497 (non-causal)"""
        s = x.permute(1, 2, 0)                                          # N x F x T
        s = F.glu(F.conv1d(s, self.p(pre + "0.weight"), self.p(pre + "0.bias")), dim=-2)
        w = self.p(pre + "2.weight")
        s = F.conv1d(s, w, self.p(pre + "2.bias"), padding=(w.shape[-1] - 1) // 2, groups=w.shape[0])
        s = F.batch_norm(s, self.p(pre + "3.running_mean"), self.p(pre + "3.running_var"), self.p(pre + "3.weight"),
                         self.p(pre + "3.bias"), False, 0.0, 1e-5)
        s = F.conv1d(self._act(s), self.p(pre + "5.weight"), self.p(pre + "5.bias"))
        return s.permute(2, 0, 1)

    def cfmr_layer(self, i: int, x, inj, kpm, amask):
        """impl.py:499-541"""
        pre = f"encoder.layers.{i}."
        mac = 0.5 if self.p(pre + "feedforward1.0.weight") is not None else 1.0
        if self.p(pre + "feedforward1.0.weight") is not None:
            if self.pre_norm:
                x = self.ffn(pre + "feedforward1.", self.ln(pre + "norm_ffn1.", x)) * mac + x
            else:
                x = self.ln(pre + "norm_ffn1.", self.ffn(pre + "feedforward1.", x) * mac + x)
        inp = self.ln(pre + "norm_attn.", x) if self.pre_norm else x
        x = x + self.attention(pre + "self_attn.", inp, inj, kpm, amask)
        if self.pre_norm:
            x = self.conv_module(pre + "convolution.", self.ln(pre + "norm_conv.", x)) + x
            return self.ffn(pre + "feedforward2.", self.ln(pre + "norm_ffn2.", x)) * mac + x
        x = self.conv_module(pre + "convolution.", self.ln(pre + "norm_attn.", x)) + x
        x = self.ln(pre + "norm_conv.", x)
        return self.ln(pre + "norm_ffn2.", self.ffn(pre + "feedforward2.", x) * mac + x)

    # ---- whole encoder (encoder.py:55-106) --------------------------------------------------------------
    def __call__(self, x: th.Tensor, lens: Optional[th.Tensor]) -> Tuple[th.Tensor, Optional[th.Tensor]]:
        proj = self.cfg.get("proj", "conv2d")
        if proj == "conv2d":
            x, lens = self.conv2d_proj(x, lens)
        elif proj == "linear":
            x, lens = self.linear_proj(x, lens)
        kpm = None if lens is None else padding_mask(lens)
        T = x.shape[1]
        inj = None
        if self.pose == "abs":
            scaled = self.cfg.get("pose_kwargs", {}).get("scaled", False)
            enc = sin_encoding(th.arange(0, T, 1.0), self.p("pose.div_term"))
            x = (x * (self.D**0.5 if scaled else 1) + enc).transpose(0, 1)
        else:
            x = x.transpose(0, 1)
            if self.pose == "rel":
                kw = self.cfg.get("pose_kwargs", {})
                lr, rr = kw.get("lradius", 128), kw.get("rradius", 128)
                pos = th.clamp(th.arange(-T + 1, T), max=rr, min=-lr)
                inj = F.embedding(pos + lr, self.p("pose.embed.weight"))
            else:
                inj = sin_encoding(th.arange(0, 2 * T - 1, 1.0), self.p("pose.div_term"))
        layer = self.cfmr_layer if self.cfg["arch"] == "cfmr" else self.xfmr_layer
        for i in range(self.cfg["num_layers"]):
            x = layer(i, x, inj, kpm, None)
        if self.p("encoder.norm.weight") is not None:
            x = self.ln("encoder.norm.", x)
        if self.p("outp.weight") is not None:
            x = F.linear(x, self.p("outp.weight"), self.p("outp.bias"))
        return x.transpose(0, 1), lens
