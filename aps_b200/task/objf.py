"""Time-domain separation objectives on the fused one-pass kernel (`aps_b200_pair_objf_fwd`).

Same names, arguments and errors as aps/task/objf.py:133-198 (sisnr_objf / snr_objf) and :244-369
(multiple_objf / permu_invarint_objf / hybrid_permu_objf).  The PIT variants build the K x K pair
matrix with ONE kernel pass over the 2K waveforms instead of K! * K calls of the pairwise objective.
Forward only (evaluation metric / loss value); CUDA tensors only.
"""
from itertools import permutations
from typing import Callable, List, Optional, Sequence

import torch as th

from .. import _lib

EPSILON = float(th.finfo(th.float32).eps)   # aps/const.py:12
MAX_SIGNALS = 4


class _Kind:
    """Tag carried by `sisnr_objf` / `snr_objf` partials so that the PIT helpers can recognise them and
    take the fused path (anything else goes through the generic pairwise loop like the reference)."""
    SISNR, SNR = 0, 1


def _signal_list(sigs: Sequence[th.Tensor]):
    sl = _lib.SignalList()
    for k, s in enumerate(sigs):
        sl.ptr[k] = s.data_ptr()
        sl.ld[k] = s.stride(0)
    sl.count = len(sigs)
    return sl


def _rows(x: th.Tensor) -> th.Tensor:
    if x.dim() == 1:
        x = x[None]
    x = x.reshape(-1, x.shape[-1])
    if x.dtype != th.float32:
        x = x.float()
    return x if x.stride(-1) == 1 else x.contiguous()


def pair_objf_matrix(est: Sequence[th.Tensor], ref: Sequence[th.Tensor], kind: int = _Kind.SISNR,
                     eps: float = EPSILON, zero_mean: bool = True, non_nagetive: bool = False,
                     snr_max: float = -1) -> th.Tensor:
    """[N, K, K] matrix  m[n, e, r] = objf(est[e][n], ref[r][n])  for K estimates / references of N x S."""
    K = len(est)
    if K != len(ref):
        raise ValueError(f"Size mismatch between #inp and #ref: {K} vs {len(ref)}")
    if not 1 <= K <= MAX_SIGNALS:
        raise RuntimeError(f"aps_b200 pair objective supports 1..{MAX_SIGNALS} signals, got {K}")
    for x in list(est) + list(ref):
        if x.shape != est[0].shape:
            raise RuntimeError("Dimention mismatch when calculate " + f"si-snr, {x.shape} vs {est[0].shape}")
    dev = _lib.require_cuda(est[0], "separated signal")
    if th.is_grad_enabled() and any(x.requires_grad for x in est):
        # the kernels compute the VALUE of the objective (the forward hot path); returning a loss without an autograd
        # graph to a training loop would silently train nothing
        raise RuntimeError("aps_b200 fused Si-SNR / SNR / PIT objectives are inference-only (no autograd graph): the "
                           "separated signals require grad; evaluate under torch.no_grad() or use the reference task "
                           "for training")
    est = [_rows(x) for x in est]
    ref = [_rows(_check_dev(s, dev)) for s in ref]
    N, S = est[0].shape
    lib = _lib.load()
    nbytes = lib.aps_b200_pair_objf_workspace_bytes(N, S, K)
    if nbytes <= 0:
        raise RuntimeError(f"aps_b200 pair objective: unsupported shape {N} x {S}")
    ws = th.empty(nbytes // 8, dtype=th.float64, device=dev)
    out = th.empty((N, K, K), dtype=th.float32, device=dev)
    d = _lib.ObjfDesc(kind=int(kind), zero_mean=int(bool(zero_mean)), non_negative=int(bool(non_nagetive)),
                      eps=float(eps), snr_max=float(snr_max))
    e_list, r_list = _signal_list(est), _signal_list(ref)
    with th.cuda.device(dev):
        _lib.check(lib.aps_b200_pair_objf_fwd(e_list, r_list, N, S, d, ws.data_ptr(), nbytes, out.data_ptr(),
                                               _lib.stream_ptr(dev)))
    return out


def _check_dev(t: th.Tensor, dev: th.device) -> th.Tensor:
    if not t.is_cuda or t.device != dev:
        raise RuntimeError(f"aps_b200 objective: reference signal on {t.device}, estimate on {dev}")
    return t


def sisnr_objf(x: th.Tensor, s: th.Tensor, eps: float = EPSILON, zero_mean: bool = True,
               non_nagetive: bool = False) -> th.Tensor:
    """Si-SNR of separated x against reference s (N x S each) -> N.  aps/task/objf.py:133-163."""
    if x.shape != s.shape:
        raise RuntimeError("Dimention mismatch when calculate " + f"si-snr, {x.shape} vs {s.shape}")
    m = pair_objf_matrix([x], [s], _Kind.SISNR, eps=eps, zero_mean=zero_mean, non_nagetive=non_nagetive)
    return m.reshape(x.shape[:-1])


def snr_objf(x: th.Tensor, s: th.Tensor, eps: float = EPSILON, snr_max: float = -1,
             non_nagetive: bool = False) -> th.Tensor:
    """SNR of separated x against reference s (N x S each) -> N.  aps/task/objf.py:166-198."""
    if x.shape != s.shape:
        raise RuntimeError("Dimention mismatch when calculate " + f"si-snr, {x.shape} vs {s.shape}")
    m = pair_objf_matrix([x], [s], _Kind.SNR, eps=eps, non_nagetive=non_nagetive, snr_max=snr_max)
    return m.reshape(x.shape[:-1])


class FusedObjf:
    """A pairwise objective `sign * kind(x, s)` that the PIT helpers may evaluate for all pairs at once."""

    def __init__(self, kind: int, sign: float = 1.0, **kwargs):
        self.kind, self.sign, self.kwargs = kind, sign, kwargs

    def __call__(self, x: th.Tensor, s: th.Tensor) -> th.Tensor:
        fn = sisnr_objf if self.kind == _Kind.SISNR else snr_objf
        return self.sign * fn(x, s, **self.kwargs)

    def matrix(self, est, ref) -> th.Tensor:
        return self.sign * pair_objf_matrix(est, ref, self.kind, **self.kwargs)


def multiple_objf(inp: List, ref: List, objf: Callable, weight: Optional[List[float]] = None,
                  transform: Optional[Callable] = None, batchmean: bool = False) -> th.Tensor:
    """Weighted sum of pairwise objectives (no permutation).  aps/task/objf.py:244-275."""
    if len(inp) != len(ref):
        raise ValueError("Size mismatch between #inp and " + f"#ref: {len(inp)} vs {len(ref)}")
    num_tasks = len(inp)
    if weight is None:
        weight = [1 / num_tasks] * num_tasks
    if len(weight) != len(inp):
        raise RuntimeError(f"Missing weight ({len(weight)}) for {len(inp)} tasks")
    if transform:
        inp = [transform(i) for i in inp]
        ref = [transform(r) for r in ref]
    loss = [objf(o, r) for o, r in zip(inp, ref)]
    loss = sum([s * l for s, l in zip(weight, loss)])
    if batchmean:
        loss = th.mean(loss)
    return loss


def permu_invarint_objf(inp: List, ref: List, objf: Callable, transform: Optional[Callable] = None,
                        batchmean: bool = False, return_permutation: bool = False):
    """Permutation-invariant objective (min over permutations).  aps/task/objf.py:278-336.

    With a `FusedObjf` the K x K pair matrix comes from one kernel pass; the permutation table is the
    same `itertools.permutations` order as the reference, so `index` matches it."""
    num_spks = len(inp)
    if num_spks != len(ref):
        raise ValueError("Size mismatch between #inp and " + f"#ref: {num_spks} vs {len(ref)}")
    if transform:
        inp = [transform(i) for i in inp]
        ref = [transform(r) for r in ref]
    if num_spks == 1:
        return objf(inp[0], ref[0])
    perms = list(permutations(range(num_spks)))
    if isinstance(objf, FusedObjf) and num_spks <= MAX_SIGNALS:
        mat = objf.matrix(inp, ref)                                      # N x K x K
        rows = th.arange(num_spks, device=mat.device)
        cols = th.tensor(perms, device=mat.device)                       # P x K
        loss_mat = mat[:, rows[None, :], cols].sum(-1).transpose(0, 1) / num_spks   # P x N
    else:
        loss_mat = th.stack([sum([objf(inp[s], ref[t]) for s, t in enumerate(p)]) / len(p) for p in perms])
    loss, index = th.min(loss_mat, dim=0)
    if batchmean:
        loss = th.mean(loss)
    if return_permutation:
        return loss, index
    return loss


def hybrid_permu_objf(out: List, ref: List, objf: Callable, transform: Optional[Callable] = None,
                      weight: Optional[List[float]] = None, permute: bool = True,
                      permu_num_spks: int = 2) -> th.Tensor:
    """Pair-wise, permutated, or permutated + pair-wise residual branches.  aps/task/objf.py:339-369."""
    num_branch = len(out)
    if num_branch != len(ref):
        raise RuntimeError(f"Got {len(ref)} references but with {num_branch} outputs")
    if permute:
        loss = permu_invarint_objf(out[:permu_num_spks], ref[:permu_num_spks], objf, transform=transform)
        if num_branch > permu_num_spks:
            num_weight = num_branch - (permu_num_spks - 1)
            if weight is None:
                weight = [1 / num_weight] * num_weight
            other_loss = multiple_objf(out[permu_num_spks:], ref[permu_num_spks:], objf, weight=weight[1:])
            loss = weight[0] * loss + other_loss
    else:
        loss = multiple_objf(out, ref, objf, weight=weight, transform=transform)
    return loss
