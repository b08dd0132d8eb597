"""CPU restatement of the reference's time-domain separation objectives — TEST INFRASTRUCTURE ONLY.

fp32 torch-on-CPU, same operation order as the reference:
  sisnr / snr      aps/task/objf.py:133-163 / :166-198
  multiple / pit   aps/task/objf.py:244-275 / :278-336
  hybrid           aps/task/objf.py:339-369
Pinned bit-for-bit to the live reference by tests/test_objf.py::test_oracle_objf_vs_live_reference.
"""
from itertools import permutations

import torch as th

EPSILON = float(th.finfo(th.float32).eps)


def _l2(m, keepdim=False):
    return th.norm(m, dim=-1, keepdim=keepdim)


def sisnr(x, s, eps=EPSILON, zero_mean=True, non_nagetive=False):
    if x.shape != s.shape:
        raise RuntimeError(f"Dimention mismatch when calculate si-snr, {x.shape} vs {s.shape}")
    if zero_mean:                                                           # objf.py:151-153
        x = x - th.mean(x, dim=-1, keepdim=True)
        s = s - th.mean(s, dim=-1, keepdim=True)
    t = th.sum(x * s, dim=-1, keepdim=True) * s / (_l2(s, keepdim=True)**2 + eps)   # objf.py:154-155
    snr_linear = _l2(t) / (_l2(x - t) + eps)                                # objf.py:157
    if non_nagetive:
        return 10 * th.log10(1 + snr_linear**2)
    return 20 * th.log10(eps + snr_linear)


def snr(x, s, eps=EPSILON, snr_max=-1, non_nagetive=False):
    if x.shape != s.shape:
        raise RuntimeError(f"Dimention mismatch when calculate si-snr, {x.shape} vs {s.shape}")
    if snr_max > 0:                                                         # objf.py:183-190
        threshold = 10**(-snr_max / 10)
        s_norm = _l2(s)**2
        x_s_norm = _l2(x - s)**2
        return 10 * th.log10(s_norm + eps) - 10 * th.log10(threshold * s_norm + x_s_norm + eps)
    snr_linear = _l2(s) / (_l2(x - s) + eps)
    if non_nagetive:
        return 10 * th.log10(1 + snr_linear**2)
    return 20 * th.log10(eps + snr_linear)


def multiple(inp, ref, objf, weight=None):
    if weight is None:
        weight = [1 / len(inp)] * len(inp)
    return sum(w * objf(o, r) for w, o, r in zip(weight, inp, ref))


def pit(inp, ref, objf, return_permutation=False):
    if len(inp) == 1:
        return objf(inp[0], ref[0])
    mat = th.stack([sum(objf(inp[s], ref[t]) for s, t in enumerate(p)) / len(p)
                    for p in permutations(range(len(inp)))])                # objf.py:318-327
    loss, index = th.min(mat, dim=0)
    return (loss, index) if return_permutation else loss


def hybrid(out, ref, objf, weight=None, permute=True, permu_num_spks=2):
    if not permute:
        return multiple(out, ref, objf, weight)
    loss = pit(out[:permu_num_spks], ref[:permu_num_spks], objf)
    if len(out) > permu_num_spks:                                           # objf.py:357-366
        nw = len(out) - (permu_num_spks - 1)
        if weight is None:
            weight = [1 / nw] * nw
        loss = weight[0] * loss + multiple(out[permu_num_spks:], ref[permu_num_spks:], objf, weight[1:])
    return loss
