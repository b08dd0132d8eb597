"""Thin Python wrappers over the dense-layer / normalisation / attention entry points of the C ABI.
Every function takes CUDA fp32 tensors, allocates the output with torch and launches on the current
stream; there is no fallback path."""
from typing import Optional

import torch as th

from . import _lib


def _epilogue(bias=None, act="none", alpha=1.0, slope=None, leaky=0.0, residual=None, beta=1.0, post=None):
    e = _lib.Epilogue()
    e.bias = _lib.ptr(bias)
    e.act = _lib.ACT[act]
    e.alpha = float(alpha)
    e.prelu_slope = _lib.ptr(slope)
    e.prelu_per_channel = int(slope is not None and slope.numel() > 1)
    e.leaky_slope = float(leaky)
    e.residual = _lib.ptr(residual)
    e.ld_residual = residual.stride(0) if residual is not None else 0
    e.beta = float(beta)
    e.post_scale, e.post_shift = (_lib.ptr(post[0]), _lib.ptr(post[1])) if post is not None else (0, 0)
    return e


import os

# GEMM engine of `linear` / `conv2d_nhwc`: "tc" (default) = tcgen05 3xTF32 tensor-core kernel whenever the shape
# allows it (K % 4 == 0, 16-byte aligned rows, M >= 64), "simt" = the exact-fp32 CUDA-core kernel everywhere
GEMM_ENGINE = os.environ.get("APS_B200_GEMM", "tc")


def tf32_split(x: th.Tensor):
    """(hi, lo) = (rn_tf32(x), rn_tf32(x - hi)) for a [rows, cols] matrix with 16-byte aligned rows."""
    dev = x.device
    buf = th.empty((2, x.shape[0], x.shape[1]), dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_tf32_split(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), buf[0].data_ptr(),
                                                   buf[1].data_ptr(), buf.stride(1), _lib.stream_ptr(dev)))
    return buf[0], buf[1]


class SplitCache:
    """Per-module cache of TF32 hi/lo splits of weight tensors.  It keeps a reference to every cached weight
    (so its address cannot be recycled under the cache) and re-splits when the tensor was modified in place."""

    def __init__(self):
        self._items = {}
        self.extra = {}          # other derived copies owned by the same module (e.g. zero-padded LSTM weights)

    def get(self, w: th.Tensor):
        key = (w.data_ptr(), tuple(w.shape), w.stride(0))
        hit = self._items.get(key)
        if hit is None or hit[1] != w._version:
            hit = self._items[key] = (w, w._version, tf32_split(w))
        return hit[2]

    def clear(self):
        self._items.clear()
        self.extra.clear()


# bench.py's kernel leg: when set to a list, every tensor-core GEMM launch appends (kind, flops, replay) where `replay()`
# re-issues exactly the same C-ABI call (same buffers): bench.py captures all of them into one CUDA graph and times the
# replay, i.e. the device time of the GEMM launches of one step without the other kernels and without host gaps.
PROFILE = None


def _tc_call(kind: str, flops: float, dev, fn, *args, keep=()):
    """Issue the C-ABI call `fn(*args)` on `dev`; remember it for bench.py when ops.PROFILE is a list (`keep`: the
    tensors behind the raw pointers in `args`, held by the replay closure so the addresses stay valid)."""
    with th.cuda.device(dev):
        _lib.check(fn(*args))
    if PROFILE is not None:
        def replay(fn=fn, args=args, dev=dev, keep=keep):
            with th.cuda.device(dev):
                _lib.check(fn(*args[:-1], _lib.stream_ptr(dev)))      # the last argument is always the stream
        PROFILE.append((kind, flops, replay))


class PackGuard:
    """Detects in-place updates of the parameters / buffers that derived weight copies were built from (BatchNorm
    folded into convolutions, GLU-interleaved rows, NHWC filter order, TF32 hi/lo splits, captured CUDA graphs).

    `stale()` compares the sum of the tensors' version counters (bumped by every in-place op: optimizer / EMA updates
    `p.mul_().add_()`, `p.copy_()`, a `load_state_dict` on a SUB-module) and their storage addresses with the values
    seen when the copies were built.  Writes through `.data` (`p.data.copy_(...)`) do not touch the version counter of
    `p`; after such a write call the owning module's `refresh_packs()`."""

    def __init__(self, module):
        self.module, self.tensors, self.key, self.calls = module, None, None, 0

    _version_of = staticmethod(__import__("operator").attrgetter("_version"))

    def _key(self):
        ts = self.tensors
        return (sum(map(self._version_of, ts)), sum(t.data_ptr() for t in ts) & 0xFFFFFFFFFFFF, len(ts))

    def mark(self):
        import itertools
        self.tensors = list(itertools.chain(self.module.parameters(), self.module.buffers()))
        self.key = self._key()

    def stale(self) -> bool:
        """Runs on every forward, AFTER the host has waited for the previous stage (the transform's NaN guard), i.e. with
        the GPU idle: 110 us for the 396 tensors of a 12-layer conformer when versions and addresses are both summed.  The
        version counters (25 us) are compared on every call; the storage addresses — which only move under `p.data = ...`,
        the case the class docstring already sends to `refresh_packs()` — on the first call and every 16th."""
        if self.tensors is None:
            return True
        self.calls += 1
        if (self.calls & 15) == 1:
            return self._key() != self.key
        return sum(map(self._version_of, self.tensors)) != self.key[0]

    def reset(self):
        self.tensors, self.key, self.calls = None, None, 0


def _tc_ok(x: th.Tensor, w: th.Tensor, M: int, K: int, N: int) -> bool:
    return (GEMM_ENGINE == "tc" and K % 4 == 0 and K >= 32 and M >= 64 and N >= 32 and x.stride(0) % 4 == 0
            and w.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0 and w.stride(1) == 1)


def _tc_conv_ok(x: th.Tensor, Cin: int, Cout: int, M: int) -> bool:
    return GEMM_ENGINE == "tc" and Cin % 32 == 0 and M >= 128 and Cout >= 4 and x.data_ptr() % 16 == 0


def rows2d(x: th.Tensor) -> th.Tensor:
    """View `x` as [rows, cols] with unit column stride (copies only if it has to)."""
    x = x.reshape(-1, x.shape[-1])
    return x if x.stride(1) == 1 else x.contiguous()


def linear(x: th.Tensor, weight: th.Tensor, bias: Optional[th.Tensor] = None, act: str = "none", alpha: float = 1.0,
           slope=None, leaky: float = 0.0, residual: Optional[th.Tensor] = None, beta: float = 1.0,
           out: Optional[th.Tensor] = None, post=None, cache: Optional["SplitCache"] = None,
           x_lo: Optional[th.Tensor] = None, want_lo: bool = False, ksplit: int = 1):
    """out[m, :] = alpha * act(x[m, :] @ weight.T + bias) + beta * residual[m, :]   (x: [M, K], weight: [N, K])

    Tensor-core extras (encoder stack): `x_lo` = the TF32 lo companion of x (see `lo_companion`) — the kernel then loads
    both operand sides by TMA; `want_lo` -> returns (out, out_lo); `ksplit` > 1 (needs x_lo, empty epilogue) -> returns
    the RAW partial sums [ksplit, M, N] that `layernorm2` reduces."""
    dev = _lib.require_cuda(x, "linear input")
    M, K = x.shape
    N = weight.shape[0]
    ncol = N // 2 if act == "glu" else N
    tc = _tc_ok(x, weight, M, K, N)
    if x_lo is not None and not (tc and x_lo.shape == x.shape and x_lo.stride() == x.stride()):
        x_lo = None
    if ksplit > 1:
        if x_lo is None:
            raise RuntimeError("ops.linear: split-K needs the tensor-core path and the lo companion of x")
        parts = th.empty((ksplit, M, N), dtype=th.float32, device=dev)
        w_hi, w_lo = cache.get(weight) if cache is not None else tf32_split(weight)
        e = _epilogue()
        _tc_call("linear", 2.0 * M * K * N, dev, _lib.load().aps_b200_linear_tc2_fwd, x.data_ptr(), x_lo.data_ptr(), M, K,
                 x.stride(0), w_hi.data_ptr(), w_lo.data_ptr(), w_hi.stride(0), N, e, parts.data_ptr(), 0, N, ksplit, M * N,
                 _lib.stream_ptr(dev), keep=(x, x_lo, w_hi, w_lo, parts))
        return parts
    if out is None:
        out = th.empty((M, ncol), dtype=th.float32, device=dev)
    e = _epilogue(bias, act, alpha, slope, leaky, residual, beta, post)
    if tc:
        w_hi, w_lo = cache.get(weight) if cache is not None else tf32_split(weight)
        lo = th.empty_like(out) if want_lo else None
        if x_lo is None and lo is None:
            _tc_call("linear", 2.0 * M * K * N, dev, _lib.load().aps_b200_linear_tc_fwd, x.data_ptr(), M, K, x.stride(0),
                     w_hi.data_ptr(), w_lo.data_ptr(), w_hi.stride(0), N, e, out.data_ptr(), out.stride(0),
                     _lib.stream_ptr(dev), keep=(x, w_hi, w_lo, out, bias, residual, slope, post))
        else:
            _tc_call("linear", 2.0 * M * K * N, dev, _lib.load().aps_b200_linear_tc2_fwd, x.data_ptr(), _lib.ptr(x_lo), M, K,
                     x.stride(0), w_hi.data_ptr(), w_lo.data_ptr(), w_hi.stride(0), N, e, out.data_ptr(), _lib.ptr(lo),
                     out.stride(0), 1, 0, _lib.stream_ptr(dev), keep=(x, x_lo, w_hi, w_lo, out, lo, bias, residual, slope, post))
        return (out, lo) if want_lo else out
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_linear_fwd(x.data_ptr(), M, K, x.stride(0), weight.data_ptr(), weight.stride(0),
                                                   N, e, out.data_ptr(), out.stride(0), _lib.stream_ptr(dev)))
    return (out, None) if want_lo else out


def conv2d_nhwc(x: th.Tensor, weight: th.Tensor, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1),
                act: str = "none", slope=None, leaky: float = 0.0, cache: Optional["SplitCache"] = None,
                want_lo: bool = False):
    """x [B, H, W, Cin] contiguous, weight [Cout, KH, KW, Cin] contiguous -> [B, OH, OW, Cout]
    (`want_lo`: -> (out, TF32 lo companion or None when the layer does not run on the tensor-core engine))."""
    dev = _lib.require_cuda(x, "conv input")
    B, H, W, Cin = x.shape
    Cout, KH, KW, _ = weight.shape
    OH = (H + 2 * padding[0] - dilation[0] * (KH - 1) - 1) // stride[0] + 1
    OW = (W + 2 * padding[1] - dilation[1] * (KW - 1) - 1) // stride[1] + 1
    if OH <= 0 or OW <= 0:
        raise RuntimeError(f"convolution output is empty for input {tuple(x.shape)}")
    out = th.empty((B, OH, OW, Cout // 2 if act == "glu" else Cout), dtype=th.float32, device=dev)
    e = _epilogue(bias, act, 1.0, slope, leaky)
    M, K = B * OH * OW, KH * KW * Cin
    if _tc_conv_ok(x, Cin, Cout, M):
        # tensor-core path: implicit im2col + TF32 split by the kernel's producer warps
        w2 = weight.view(Cout, K)
        w_hi, w_lo = cache.get(w2) if cache is not None else tf32_split(w2)
        lo = th.empty_like(out) if (want_lo and act != "glu" and Cout % 4 == 0) else None
        _tc_call("conv2d", 2.0 * M * K * Cout, dev, _lib.load().aps_b200_conv2d_nhwc_tc2_fwd, x.data_ptr(), B, H, W, Cin,
                 w_hi.data_ptr(), w_lo.data_ptr(), Cout, KH, KW, stride[0], stride[1], padding[0], padding[1], dilation[0],
                 dilation[1], e, out.data_ptr(), _lib.ptr(lo), _lib.stream_ptr(dev), keep=(x, w_hi, w_lo, out, lo, bias, slope))
        return (out, lo) if want_lo else out
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_conv2d_nhwc_fwd(x.data_ptr(), B, H, W, Cin, weight.data_ptr(), Cout, KH, KW,
                                                        stride[0], stride[1], padding[0], padding[1], dilation[0],
                                                        dilation[1], e, out.data_ptr(), _lib.stream_ptr(dev)))
    return (out, None) if want_lo else out


def cat_complex(a: th.Tensor, b: th.Tensor) -> th.Tensor:
    """Channel concat of two stacked-complex NHWC tensors: [re_a | re_b | im_a | im_b]."""
    ca, cb = a.shape[-1] // 2, b.shape[-1] // 2
    return th.cat([a[..., :ca], b[..., :cb], a[..., ca:], b[..., cb:]], -1)


def conv_transpose2d_nhwc(x: th.Tensor, weight: th.Tensor, bias=None, stride=(1, 1), padding=(0, 0),
                          output_padding=(0, 0), act: str = "none", leaky: float = 0.0,
                          cache: Optional["SplitCache"] = None, skip: Optional[th.Tensor] = None) -> th.Tensor:
    """x [B, H, W, Cin], weight [Cout, KH, KW, Cin] -> [B, OH, OW, Cout] (transposed convolution).
    `skip` (same shape as x): the input is cat_complex(x, skip); the tensor-core engine reads both in place."""
    dev = _lib.require_cuda(x, "conv input")
    Cout, KH, KW, _ = weight.shape
    B, H, W, Cx = x.shape
    OH = (H - 1) * stride[0] - 2 * padding[0] + KH + output_padding[0]
    OW = (W - 1) * stride[1] - 2 * padding[1] + KW + output_padding[1]
    Cskip = Cx if skip is not None else 0
    if (Cout <= 8 and KH * KW * ((Cx + Cskip + 63) // 64) * 64 * 32 <= 48 * 1024 and Cx % (8 if skip is not None else 4) == 0
            and x.is_contiguous() and x.data_ptr() % 16 == 0
            and (skip is None or (skip.shape == x.shape and skip.is_contiguous() and skip.data_ptr() % 16 == 0))):
        # a handful of output channels (the last decoder layer): dedicated kernel, skip tensor read in place
        out = th.empty((B, OH, OW, Cout), dtype=th.float32, device=dev)
        e = _epilogue(bias, act, 1.0, None, leaky)
        with th.cuda.device(dev):
            _lib.check(_lib.load().aps_b200_conv_transpose2d_nhwc_narrow_fwd(
                x.data_ptr(), _lib.ptr(skip), B, H, W, Cx + Cskip, weight.data_ptr(), Cout, KH, KW, stride[0], stride[1],
                padding[0], padding[1], output_padding[0], output_padding[1], e, out.data_ptr(), _lib.stream_ptr(dev)))
        return out
    fused_skip = None
    if skip is not None:
        if (skip.shape == x.shape and Cx % 64 == 0 and stride[1] == 1 and x.is_contiguous() and skip.is_contiguous()
                and skip.data_ptr() % 16 == 0 and _tc_conv_ok(x, 2 * Cx, Cout, B * OH * OW)):
            fused_skip = skip
        else:
            x = cat_complex(x, skip)
    Cin = 2 * Cx if fused_skip is not None else x.shape[-1]
    out = th.empty((B, OH, OW, Cout), dtype=th.float32, device=dev)
    e = _epilogue(bias, act, 1.0, None, leaky)
    if _tc_conv_ok(x, Cin, Cout, B * OH * OW) and stride[1] == 1:     # the engine's transposed gather needs stride_w == 1
        w2 = weight.view(Cout, KH * KW * Cin)
        w_hi, w_lo = cache.get(w2) if cache is not None else tf32_split(w2)
        # FLOPs actually executed: with stride 2 along H the class-major row order skips the taps that are identically
        # zero for a tile (half of the reference's conv_transpose2d arithmetic)
        flops = 2.0 * B * OH * OW * KH * KW * Cin * Cout / (stride[0] if stride[0] in (2, 3, 4) else 1)
        _tc_call("conv_transpose2d", flops, dev, _lib.load().aps_b200_conv_transpose2d_nhwc_tc_fwd, x.data_ptr(),
                 _lib.ptr(fused_skip), B, H, W, Cin, w_hi.data_ptr(), w_lo.data_ptr(), Cout, KH, KW, stride[0], stride[1],
                 padding[0], padding[1], output_padding[0], output_padding[1], e, out.data_ptr(), _lib.stream_ptr(dev),
                 keep=(x, fused_skip, w_hi, w_lo, out, bias))
        return out
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_conv_transpose2d_nhwc_fwd(
            x.data_ptr(), B, H, W, Cin, weight.data_ptr(), Cout, KH, KW, stride[0], stride[1], padding[0], padding[1],
            output_padding[0], output_padding[1], e, out.data_ptr(), _lib.stream_ptr(dev)))
    return out


def cmask(mask: th.Tensor, col_real: int, col_imag: int, stft: Optional[th.Tensor], act: str, eps: float,
          apply: bool) -> th.Tensor:
    """Complex ratio mask from channels (col_real, col_imag) of mask [..., C]; returns [..., 2]."""
    dev = _lib.require_cuda(mask, "mask")
    pos = mask.numel() // mask.shape[-1]
    out = th.empty(mask.shape[:-1] + (2,), dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_cmask_fwd(mask.data_ptr(), mask.shape[-1], col_real, col_imag, _lib.ptr(stft),
                                                  pos, _lib.ACT[act], float(eps), int(apply), out.data_ptr(),
                                                  _lib.stream_ptr(dev)))
    return out


def layernorm(x: th.Tensor, gamma, beta, eps: float = 1e-5, residual: Optional[th.Tensor] = None,
              alpha: float = 1.0) -> th.Tensor:
    """LayerNorm(alpha * x + residual) over the last axis of [M, D] rows."""
    dev = _lib.require_cuda(x, "layernorm input")
    M, D = x.shape
    out = th.empty((M, D), dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_layernorm_fwd(x.data_ptr(), x.stride(0), _lib.ptr(residual),
                                                      residual.stride(0) if residual is not None else 0, float(alpha),
                                                      _lib.ptr(gamma), _lib.ptr(beta), float(eps), M, D, out.data_ptr(),
                                                      out.stride(0), _lib.stream_ptr(dev)))
    return out


def layernorm2(x: th.Tensor, gamma, beta, eps: float = 1e-5, bias=None, residual: Optional[th.Tensor] = None,
               alpha: float = 1.0, normalize: bool = True, want_lo: bool = True):
    """v = alpha * (sum_p x[p] + bias) + residual; y = LN(v) * gamma + beta (or v when not `normalize`).
    x: [M, D] or the split-K partial sums [P, M, D] of `linear(..., ksplit=P)`; D % 128 == 0, D <= 1024.
    -> (y, TF32 lo companion of y or None)."""
    dev = _lib.require_cuda(x, "layernorm input")
    parts = x.shape[0] if x.dim() == 3 else 1
    M, D = x.shape[-2], x.shape[-1]
    out = th.empty((M, D), dtype=th.float32, device=dev)
    lo = th.empty_like(out) if want_lo else None
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_layernorm2_fwd(x.data_ptr(), x.stride(-2), parts, x.stride(0) if x.dim() == 3 else 0,
                                                       _lib.ptr(bias), _lib.ptr(residual),
                                                       residual.stride(0) if residual is not None else 0, float(alpha),
                                                       _lib.ptr(gamma), _lib.ptr(beta), float(eps), int(normalize), M, D,
                                                       out.data_ptr(), _lib.ptr(lo), out.stride(0), _lib.stream_ptr(dev)))
    return out, lo


def lo_companion(x: th.Tensor) -> th.Tensor:
    """rn_tf32(x - trunc_tf32(x)) for a tensor that no kernel epilogue produced (e.g. the input features): what a
    TMA-fed tensor-core GEMM needs next to the raw x.  Plain device tensor ops (integer view arithmetic)."""
    xi = x.contiguous().view(th.int32)
    lo = x - (xi & -8192).view(th.float32)
    return ((lo.view(th.int32) + 4096) & -8192).view(th.float32)


def layernorm2_ok(D: int) -> bool:
    return D % 128 == 0 and D <= 1024


def utt_norm(x: th.Tensor, N: int, T: int, gamma=None, beta=None, eps: float = 1e-5, per_channel: bool = False,
             relu: bool = False, stride_n: Optional[int] = None, stride_t: int = 1, inplace: bool = False) -> th.Tensor:
    """Per-utterance normalisation over time of token rows [N*T, C] (row(n, t) = n*stride_n + t*stride_t):
    GroupNorm(1, C) / gLN statistics over (C, T), or GroupNorm(C, C) statistics over T when `per_channel`."""
    dev = _lib.require_cuda(x, "normalisation input")
    C = x.shape[1]
    lib = _lib.load()
    nbytes = lib.aps_b200_utt_norm_workspace_bytes(N, T, C)
    ws = th.empty(nbytes // 8, dtype=th.float64, device=dev)
    out = x if inplace else th.empty_like(x)
    with th.cuda.device(dev):
        _lib.check(lib.aps_b200_utt_norm_fwd(x.data_ptr(), x.stride(0), N, T, C, T if stride_n is None else stride_n,
                                             stride_t, int(per_channel), _lib.ptr(gamma), _lib.ptr(beta), float(eps),
                                             int(relu), ws.data_ptr(), nbytes, out.data_ptr(), out.stride(0),
                                             _lib.stream_ptr(dev)))
    return out


def lstm_multi(xs, mods, caches=None):
    """torch.nn.LSTM forward (batch_first, zero initial state, inference) of several modules of identical shape on
    identically shaped inputs x [N, T, In]: per layer one tensor-core GEMM per module and direction for the input
    projections of all frames, then ONE fused recurrence launch per frame that advances all modules and directions
    together (aps_b200_lstm_group_fwd).  Returns a list of [N, T, H * directions]."""
    import ctypes
    m0 = mods[0]
    dev = _lib.require_cuda(xs[0], "LSTM input")
    for m in mods:
        if not m.batch_first or getattr(m, "proj_size", 0):
            raise RuntimeError("ops.lstm: only batch_first nn.LSTM without projections is implemented")
        if m.training and m.dropout > 0 and m.num_layers > 1:
            raise RuntimeError("ops.lstm: inter-layer dropout (training mode) is not implemented")
        if (m.hidden_size, m.num_layers, m.bidirectional, m.input_size, m.bias) != (
                m0.hidden_size, m0.num_layers, m0.bidirectional, m0.input_size, m0.bias):
            raise RuntimeError("ops.lstm_multi: the modules must have the same shape")
    N, T, _ = xs[0].shape
    H, dirs = m0.hidden_size, 2 if m0.bidirectional else 1
    # The kernel moves 16-byte rows: a hidden size that is not a multiple of 4 runs ZERO PADDED to Hp (padded units have
    # zero weights and biases: i = f = o = 1/2, g = 0, so their cell and output stay exactly 0 and feed nothing) and the
    # result is sliced back — no library fallback.
    Hp = (H + 3) // 4 * 4

    def gate_rows(w):                                            # [4H, X] -> [4Hp, X]
        return w if Hp == H else th.nn.functional.pad(w.reshape(4, H, -1), (0, 0, 0, Hp - H)).reshape(4 * Hp, -1)

    def in_cols(w, layer):                                       # columns of w_ih for layers fed by a padded output
        if Hp == H or layer == 0:
            return w
        return th.nn.functional.pad(w.reshape(w.shape[0], dirs, H), (0, Hp - H)).reshape(w.shape[0], dirs * Hp)

    if any(x.shape != xs[0].shape for x in xs):
        raise RuntimeError("ops.lstm_multi: the inputs must have the same shape")
    caches = caches or [None] * len(mods)
    lib = _lib.load()
    max_groups = 4                                               # APS_B200_LSTM_MAX_GROUPS
    inps = [x.contiguous().float() for x in xs]
    # Recurrence on the tensor-core engine (aps_b200_lstm_group_tc_fwd): unidirectional layers whose hidden size is a
    # multiple of 32, enough rows to fill 128-row tiles; APS_B200_LSTM=simt keeps the fp32-FMA recurrence kernel
    tc_rec = (GEMM_ENGINE == "tc" and dirs == 1 and Hp == H and H % 32 == 0 and N >= 64 and len(mods) <= max_groups
              and os.environ.get("APS_B200_LSTM", "tc") != "simt")
    if tc_rec:
        G = len(mods)
        Np = (N + 127) // 128 * 128
        for layer in range(m0.num_layers):
            sfx = f"_l{layer}"
            xg = th.empty(G, Np, T, 4 * H, dtype=th.float32, device=dev)
            w_hhs = []
            for g, (inp, mod, cache) in enumerate(zip(inps, mods, caches)):
                w_ih = getattr(mod, "weight_ih" + sfx).detach()
                bias = (getattr(mod, "bias_ih" + sfx).detach() + getattr(mod, "bias_hh" + sfx).detach()) if mod.bias else None
                linear(inp.view(N * T, -1), w_ih, bias, cache=cache if isinstance(cache, SplitCache) else None,
                       out=xg[g, :N].view(N * T, 4 * H))
                w_hhs.append(getattr(mod, "weight_hh" + sfx).detach())
            # stacked, split recurrent weights (cached in the first module's store, keyed by the parameters' versions)
            store = caches[0].extra if isinstance(caches[0], SplitCache) else None
            ver = tuple((w._version, w.data_ptr()) for w in w_hhs)
            packed = store.get(("lstm_tc", layer)) if store is not None else None
            if packed is None or packed[0] != ver:
                packed = (ver,) + tf32_split(th.cat([w.float() for w in w_hhs], 0).contiguous()) + (w_hhs,)   # keeps the weights alive
                if store is not None:
                    store[("lstm_tc", layer)] = packed
            w_hi, w_lo = packed[1], packed[2]
            ys = [th.empty(N, T, H, dtype=th.float32, device=dev) for _ in mods]
            work = th.empty(G * Np * H * 21, dtype=th.float32, device=dev)
            yarr = (ctypes.c_void_p * G)(*[y.data_ptr() for y in ys])
            with th.cuda.device(dev):
                _lib.check(lib.aps_b200_lstm_group_tc_fwd(xg.data_ptr(), N, Np, T, H, w_hi.data_ptr(), w_lo.data_ptr(), yarr,
                                                          H, G, work.data_ptr(), _lib.stream_ptr(dev)))
            inps = ys
        return inps
    for layer in range(m0.num_layers):
        ys = [th.empty(N, T, Hp * dirs, dtype=th.float32, device=dev) for _ in mods]
        jobs = []                                                # (xg, w_hh, cell, y pointer, reverse)
        for inp, mod, cache, y in zip(inps, mods, caches, ys):
            for d in range(dirs):
                sfx = f"_l{layer}" + ("_reverse" if d else "")
                w_ih = getattr(mod, "weight_ih" + sfx).detach()
                w_hh = getattr(mod, "weight_hh" + sfx).detach()
                bias = (getattr(mod, "bias_ih" + sfx).detach() + getattr(mod, "bias_hh" + sfx).detach()) if mod.bias else None
                if Hp != H:
                    key = ("lstm_pad", layer, d)
                    store = cache.extra if isinstance(cache, SplitCache) else None
                    packed = store.get(key) if store is not None else None
                    src_ver = (w_ih._version, w_hh._version, w_ih.data_ptr(), w_hh.data_ptr())
                    if packed is None or packed[0] != src_ver:
                        pw_ih = gate_rows(in_cols(w_ih, layer)).contiguous()
                        pw_hh = gate_rows(th.nn.functional.pad(w_hh, (0, Hp - H))).contiguous()
                        pb = gate_rows(bias[:, None]).reshape(-1).contiguous() if bias is not None else None
                        packed = (src_ver, pw_ih, pw_hh, pb)
                        if store is not None:
                            store[key] = packed
                    _, w_ih, w_hh, bias = packed
                else:
                    w_hh = w_hh.contiguous()
                split_cache = cache if isinstance(cache, SplitCache) else None
                xg = linear(inp.view(N * T, -1), w_ih, bias, cache=split_cache)
                jobs.append((xg, w_hh, th.empty(N, Hp, dtype=th.float32, device=dev), y.data_ptr() + 4 * Hp * d, d))
        for i in range(0, len(jobs), max_groups):
            part = jobs[i:i + max_groups]
            n = len(part)
            arr = lambda vals: (ctypes.c_void_p * n)(*vals)
            mask = sum(1 << g for g, j in enumerate(part) if j[4])
            with th.cuda.device(dev):
                _lib.check(lib.aps_b200_lstm_group_fwd(arr([j[0].data_ptr() for j in part]), part[0][0].stride(0), N, T, Hp,
                                                       arr([j[1].data_ptr() for j in part]), mask,
                                                       arr([j[2].data_ptr() for j in part]), arr([j[3] for j in part]),
                                                       Hp * dirs, n, _lib.stream_ptr(dev)))
        inps = ys
    if Hp != H:
        inps = [y.view(N, T, dirs, Hp)[..., :H].reshape(N, T, dirs * H) for y in inps]
    return inps


def lstm(x: th.Tensor, lstm_mod, cache: Optional[dict] = None) -> th.Tensor:
    """One nn.LSTM (see lstm_multi)."""
    return lstm_multi([x], [lstm_mod], [cache])[0]


def dwconv1d(x: th.Tensor, N: int, T: int, weight_kd: th.Tensor, bias, dilation: int = 1, left_pad: int = 0,
             stride_n: Optional[int] = None, stride_t: int = 1, act: str = "none", slope=None,
             residual=None, post=None, want_lo: bool = False, lens: Optional[th.Tensor] = None):
    """Depthwise conv over time on token rows [N*T, D] (row(n, t) = n*stride_n + t*stride_t);
    `want_lo` -> (out, TF32 lo companion); `lens` (device int64 [N]): frames t >= lens[n] read as zero."""
    dev = _lib.require_cuda(x, "dwconv input")
    D = x.shape[1]
    Kw = weight_kd.shape[0]
    out = th.empty_like(x)
    lo = th.empty_like(x) if want_lo else None
    e = _epilogue(None, act, 1.0, slope, 0.0, residual, 1.0, post)
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_dwconv1d2_fwd(x.data_ptr(), x.stride(0), N, T, D,
                                                      T if stride_n is None else stride_n, stride_t,
                                                      weight_kd.data_ptr(), _lib.ptr(bias), Kw, dilation, left_pad, e,
                                                      out.data_ptr(), _lib.ptr(lo), out.stride(0), _lib.ptr(lens),
                                                      _lib.stream_ptr(dev)))
    return (out, lo) if want_lo else out


def mhsa(qkv: th.Tensor, N: int, L: int, H: int, mode: int = 0, pos: Optional[th.Tensor] = None,
         rel_u=None, rel_v=None, kpm: Optional[th.Tensor] = None, kpm_fill: float = float("-inf"),
         attn_mask: Optional[th.Tensor] = None, qpos_is_value: bool = False, want_lo: bool = False):
    """Self-attention on the packed projection qkv [N*L, 3E] (rows batch-major) -> context [N*L, E]
    (`want_lo` -> (context, TF32 lo companion))."""
    dev = _lib.require_cuda(qkv, "attention input")
    E = qkv.shape[1] // 3
    dh = E // H
    out = th.empty((N * L, E), dtype=th.float32, device=dev)
    d = _lib.AttnDesc()
    base, ld = qkv.data_ptr(), qkv.stride(0)
    d.q, d.k, d.v = base, base + 4 * E, base + 8 * E
    d.qpos = d.v if qpos_is_value else d.q
    d.ld_q = d.ld_k = d.ld_v = d.ld_qpos = ld
    d.stride_n, d.stride_t = L, 1
    d.batch, d.length, d.heads, d.head_dim = N, L, H, dh
    d.mode = mode
    d.pos = _lib.ptr(pos)
    d.ld_pos = pos.stride(0) if pos is not None else 0
    d.rel_u, d.rel_v = _lib.ptr(rel_u), _lib.ptr(rel_v)
    d.key_padding_mask = _lib.ptr(kpm)
    d.padding_fill = kpm_fill
    d.attn_mask = _lib.ptr(attn_mask)
    d.scale = 1.0 / dh**0.5
    lo = th.empty_like(out) if want_lo else None
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_mhsa2_fwd(d, out.data_ptr(), _lib.ptr(lo), out.stride(0), _lib.stream_ptr(dev)))
    return (out, lo) if want_lo else out


# ------------------------------------------------------------------------------------ feature-chain tokens (csrc/featops.cu)
def project_rows(x: th.Tensor, weight: th.Tensor) -> th.Tensor:
    """x [..., K] @ weight[N, K].T on this package's GEMM kernels (mel projection of unfused chains, DCT): no library GEMM."""
    rows = rows2d(x.detach().float())
    w = weight.detach().float().contiguous()
    return linear(rows, w).view(*x.shape[:-1], w.shape[0])


def specaug_apply(x: th.Tensor, mask: th.Tensor, mask_zero: bool = True) -> th.Tensor:
    """x N x (C) x T x F, mask N x T x F (0 / 1): x * mask, or mean(x) where the mask is 0 (asr.py:678-683)."""
    dev = _lib.require_cuda(x, "SpecAugment input")
    x = x.detach().float().contiguous()
    mask = mask.detach().float().contiguous()
    N, C = x.shape[0], (x.shape[1] if x.dim() == 4 else 1)
    T, F = x.shape[-2], x.shape[-1]
    if mask.shape != (N, T, F) or mask.device != dev:
        raise RuntimeError(f"SpecAugment mask {tuple(mask.shape)} on {mask.device} does not fit features {tuple(x.shape)} on {dev}")
    lib = _lib.load()
    out = th.empty_like(x)
    nbytes = 0 if mask_zero else lib.aps_b200_specaug_workspace_bytes(x.numel())
    ws = th.empty(max(nbytes // 8, 1), dtype=th.float64, device=dev)
    with th.cuda.device(dev):
        _lib.check(lib.aps_b200_specaug_apply(x.data_ptr(), N, C, T, F, mask.data_ptr(), int(bool(mask_zero)), ws.data_ptr(),
                                              nbytes, out.data_ptr(), _lib.stream_ptr(dev)))
    return out


def splice(x: th.Tensor, lctx: int, rctx: int, subsampling: int = 1) -> th.Tensor:
    """N x ... x T x F -> N x ... x (T // subsampling) x (lctx + rctx + 1) * F, edge frames clamped (asr.py:687-728)."""
    dev = _lib.require_cuda(x, "splice input")
    x = x.detach().float().contiguous()
    T, F = x.shape[-2], x.shape[-1]
    rows = x.numel() // (T * F)
    To = T if subsampling == 1 else T // subsampling
    out = th.empty(x.shape[:-2] + (To, (lctx + rctx + 1) * F), dtype=th.float32, device=dev)
    if out.numel():
        with th.cuda.device(dev):
            _lib.check(_lib.load().aps_b200_splice_fwd(x.data_ptr(), rows, T, F, lctx, rctx, subsampling, out.data_ptr(),
                                                       _lib.stream_ptr(dev)))
    return out


def delta(x: th.Tensor, scale: th.Tensor, order: int, as_channel: bool = False) -> th.Tensor:
    """x N x (C) x T x F -> cat([x, d1, .., d_order], -1) or stack(.., 1) (asr.py:731-781); scale: the 2*ctx+1 taps."""
    dev = _lib.require_cuda(x, "delta input")
    x = x.detach().float().contiguous()
    T, F = x.shape[-2], x.shape[-1]
    rows = x.numel() // (T * F)
    ctx = (scale.numel() - 1) // 2
    sc = scale.detach().to(device=dev, dtype=th.float32).contiguous()       # the layer may still live on the host
    K = order + 1
    lib = _lib.load()
    if as_channel:
        if x.dim() != 3:
            raise RuntimeError(f"delta_as_channel expects N x T x F features, got {x.dim()}D")
        out = th.empty((x.shape[0], K, T, F), dtype=th.float32, device=dev)
        rs, ts, slot = K * T * F, F, T * F                    # element (row, k, t, f)
    else:
        out = th.empty(x.shape[:-1] + (K * F,), dtype=th.float32, device=dev)
        rs, ts, slot = T * K * F, K * F, F                    # element (row, t, k*F + f)
    view = out.view(-1)
    # slot 0 is the input itself (scale = [1] over a zero context)
    one = th.ones(1, dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(lib.aps_b200_delta_fwd(x.data_ptr(), T * F, F, rows, T, F, 0, one.data_ptr(), view.data_ptr(), rs, ts, st))
        for k in range(1, K):
            src = view.data_ptr() + 4 * (k - 1) * slot
            dst = view.data_ptr() + 4 * k * slot
            _lib.check(lib.aps_b200_delta_fwd(src, rs, ts, rows, T, F, ctx, sc.data_ptr(), dst, rs, ts, st))
    return out


def speed_perturb(wav: th.Tensor, choice: th.Tensor, weights) -> th.Tensor:
    """wav N x S; utterance n is resampled with weights[choice[n]] ([dst, src, K] each) or kept when choice[n] ==
    len(weights); the result is zero padded to the longest utterance (asr.py:168-195)."""
    import ctypes
    dev = _lib.require_cuda(wav, "speed perturb input")
    wav = wav.detach().float()
    if wav.stride(-1) != 1:
        wav = wav.contiguous()
    N, S = wav.shape
    ws = [w.detach().float().contiguous() for w in weights]
    nf = len(ws)
    lens = [S if c == nf else (S // ws[c].shape[1]) * ws[c].shape[0] for c in choice.tolist()]
    for c in set(choice.tolist()):
        if c != nf and S // ws[c].shape[1] == 0:
            raise RuntimeError(f"Input wav is too short to be perturbed, length = {S}")
    ld_out = max(lens)
    out = th.empty((N, ld_out), dtype=th.float32, device=dev)
    ch = choice.to(device=dev, dtype=th.int32).contiguous()
    arr = lambda vals: (ctypes.c_int32 * max(nf, 1))(*vals) if nf else None
    ptrs = (ctypes.c_void_p * max(nf, 1))(*[w.data_ptr() for w in ws]) if nf else None
    with th.cuda.device(dev):
        _lib.check(_lib.load().aps_b200_speed_perturb_fwd(wav.data_ptr(), N, S, wav.stride(0), ch.data_ptr(), nf, ptrs,
                                                          arr([w.shape[0] for w in ws]), arr([w.shape[1] for w in ws]),
                                                          arr([w.shape[2] for w in ws]), out.data_ptr(), ld_out,
                                                          _lib.stream_ptr(dev)))
    return out
