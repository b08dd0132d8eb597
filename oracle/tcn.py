"""Oracle (CPU, fp32) for row a19 of SURVEY.md §8: the frequency-domain Conv-TasNet mask estimator
(`sse@freq_tcn`), evaluated functionally from a reference `state_dict` in eval mode.
TEST INFRASTRUCTURE — see oracle/__init__.py.  Follows /root/reference/aps/sse/bss/tcn.py:112-226
(Conv1dBlock, Conv1dRepeat), :361-414 (FreqConvTasNet._tf_mask) and aps/sse/base.py:112-156
(MaskNonLinear)."""
from typing import Dict, List

import torch as th
import torch.nn.functional as F


def _norm(sd, pre, x, norm):
    if norm == "BN":
        return F.batch_norm(x, sd[pre + "running_mean"], sd[pre + "running_var"], sd[pre + "weight"], sd[pre + "bias"],
                            False, 0.0, 1e-5)
    if norm == "cLN":
        return F.group_norm(x, 1, sd[pre + "weight"], sd[pre + "bias"])
    if norm == "IN":
        return F.group_norm(x, x.shape[1], sd[pre + "weight"], sd[pre + "bias"])
    mean = x.mean((1, 2), keepdim=True)                       # gLN, tcn.py:33-72
    var = ((x - mean)**2).mean((1, 2), keepdim=True)
    return sd[pre + "gamma"] * (x - mean) / th.sqrt(var + 1e-5) + sd[pre + "beta"]


def _scale_linear(sd, pre, x):
    """ScaleLinear (tcn.py:91-109): 1x1 conv times an optional learnt scalar."""
    y = F.conv1d(x, sd[pre + "weight"], sd.get(pre + "bias"))
    return y * sd[pre + "scale"] if pre + "scale" in sd else y


def conv1d_block(sd: Dict[str, th.Tensor], pre: str, x: th.Tensor, dilation: int, norm: str, causal: bool) -> th.Tensor:
    """tcn.py:112-159"""
    K = sd[pre + "dconv.weight"].shape[-1]
    pad = dilation * (K - 1)
    y = _scale_linear(sd, pre + "conv1.", x)
    y = _norm(sd, pre + "norm1.1.", F.prelu(y, sd[pre + "norm1.0.weight"]), norm)
    y = F.conv1d(y, sd[pre + "dconv.weight"], sd[pre + "dconv.bias"], padding=pad if causal else pad // 2,
                 dilation=dilation, groups=y.shape[1])
    if causal:
        y = y[..., :-pad]
    y = _norm(sd, pre + "norm2.1.", F.prelu(y, sd[pre + "norm2.0.weight"]), norm)
    return _scale_linear(sd, pre + "conv2.", y) + x


def tf_mask(sd: Dict[str, th.Tensor], feats: th.Tensor, num_repeats: int = 3, num_blocks: int = 6,
            num_spks: int = 2, norm: str = "BN", non_linear: str = "relu", causal: bool = False,
            skip_residual: bool = False) -> List[th.Tensor]:
    """feats N x T x F -> [N x F x T] * num_spks.  tcn.py:403-414 (+ :162-226 for the repeat stack)."""
    x = F.conv1d(feats.transpose(1, 2), sd["proj.1.weight"], sd["proj.1.bias"])
    outs, skip = [x], 0
    for r in range(num_repeats):
        if skip_residual:                                    # tcn.py:203-224; the reference adds IN PLACE, so the
            for i in range(r):                               # stored output of the previous repeat changes too
                x = x + _scale_linear(sd, f"conv.skip_linear.{skip + i}.", outs[i])
            outs[r] = x
            skip += r
        for b in range(num_blocks):
            x = conv1d_block(sd, f"conv.repeat.{r}.{b}.", x, 2**b, norm, causal)
        outs.append(x)
    m = F.conv1d(F.prelu(x, sd["mask.0.weight"]), sd["mask.1.weight"], sd["mask.1.bias"])
    m = {"relu": th.relu, "sigmoid": th.sigmoid}[non_linear](m)
    return list(th.chunk(m, num_spks, 1))


def time_tcn_forward(sd: Dict[str, th.Tensor], mix: th.Tensor, L: int = 20, num_repeats: int = 4, num_blocks: int = 8,
                     num_spks: int = 2, norm: str = "BN", non_linear: str = "relu", causal: bool = False,
                     skip_residual: bool = False) -> List[th.Tensor]:
    """TimeConvTasNet.forward (tcn.py:326-358): mix N x S -> [N x S', ...]."""
    w = th.relu(F.conv1d(mix[:, None], sd["encoder.weight"], sd["encoder.bias"], stride=L // 2))     # tcn.py:335
    y = F.group_norm(w, 1, sd["ln.weight"], sd["ln.bias"])                                          # cLN, tcn.py:262
    x = F.conv1d(y, sd["proj.weight"], sd["proj.bias"])
    outs, skip = [x], 0
    for r in range(num_repeats):
        if skip_residual:
            for i in range(r):
                x = x + _scale_linear(sd, f"conv.skip_linear.{skip + i}.", outs[i])
            outs[r] = x
            skip += r
        for b in range(num_blocks):
            x = conv1d_block(sd, f"conv.repeat.{r}.{b}.", x, 2**b, norm, causal)
        outs.append(x)
    e = F.conv1d(F.prelu(x, sd["mask.0.weight"]), sd["mask.1.weight"], sd["mask.1.bias"])
    m = th.stack(th.chunk(e, num_spks, 1), 0)                                                        # tcn.py:344-346
    m = {"relu": th.relu, "sigmoid": th.sigmoid, "softmax": lambda t: th.softmax(t, 0)}[non_linear](m)
    return [F.conv_transpose1d(w * m[n], sd["decoder.weight"], sd["decoder.bias"], stride=L // 2)[:, 0]
            for n in range(num_spks)]
