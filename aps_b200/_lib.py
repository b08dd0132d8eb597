"""ctypes binding of libaps_b200.so (the C ABI declared in include/aps_b200.h).

There is NO fallback: if the library is missing or a tensor is not on a CUDA device the
callers raise RuntimeError.  The library is built in-tree by `python -m aps_b200.build`
(or `__graft_entry__.build()`).
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from typing import Dict, Tuple

import numpy as np
import torch as th

_HERE = os.path.dirname(os.path.abspath(__file__))
# APS_B200_LIB selects a tuning build of the same ABI (see aps_b200/build.py); default: the in-tree library
LIB_PATH = os.environ.get("APS_B200_LIB") or os.path.join(_HERE, "libaps_b200.so")
ABI_VERSION = 6


class StftDesc(Structure):
    """mirror of `aps_b200_stft_desc`"""
    _fields_ = [("nfft", c_int32), ("frame_width", c_int32), ("hop", c_int32), ("center_pad", c_int32),
                ("rescale", c_int32), ("utt_preemph", c_float), ("frame_preemph", c_float),
                ("frame_one_minus", c_float), ("scale", c_float), ("window", c_void_p),
                ("twiddles", c_void_p)]


class FeatDesc(Structure):
    """mirror of `aps_b200_feat_desc`"""
    _fields_ = [("power", c_int32), ("num_mels", c_int32), ("mel_start", c_void_p), ("mel_len", c_void_p),
                ("mel_weight", c_void_p), ("mel_stride", c_int32), ("log_mode", c_int32),
                ("log_eps", c_float), ("log_lower_bound", c_float), ("cmvn_mode", c_int32),
                ("norm_mean", c_int32), ("norm_var", c_int32), ("cmvn_eps", c_float), ("gmean", c_void_p),
                ("gstd", c_void_p), ("nan_count", c_void_p), ("aug_mask", c_void_p)]


class Epilogue(Structure):
    """mirror of `aps_b200_epilogue`"""
    _fields_ = [("bias", c_void_p), ("act", c_int32), ("alpha", c_float), ("prelu_slope", c_void_p),
                ("prelu_per_channel", c_int32), ("leaky_slope", c_float), ("residual", c_void_p),
                ("ld_residual", c_int64), ("beta", c_float), ("post_scale", c_void_p), ("post_shift", c_void_p)]


class AttnDesc(Structure):
    """mirror of `aps_b200_attn_desc`"""
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("qpos", c_void_p), ("ld_q", c_int64),
                ("ld_k", c_int64), ("ld_v", c_int64), ("ld_qpos", c_int64), ("stride_n", c_int64),
                ("stride_t", c_int64), ("batch", c_int64), ("length", c_int64), ("heads", c_int64),
                ("head_dim", c_int64), ("mode", c_int32), ("pos", c_void_p), ("ld_pos", c_int64),
                ("rel_u", c_void_p), ("rel_v", c_void_p), ("key_padding_mask", c_void_p),
                ("padding_fill", c_float), ("attn_mask", c_void_p), ("scale", c_float)]


class SignalList(Structure):
    """mirror of `aps_b200_signal_list`"""
    _fields_ = [("ptr", c_void_p * 4), ("ld", c_int64 * 4), ("count", c_int32)]


class ObjfDesc(Structure):
    """mirror of `aps_b200_objf_desc`"""
    _fields_ = [("kind", c_int32), ("zero_mean", c_int32), ("non_negative", c_int32), ("eps", c_float),
                ("snr_max", c_float)]


ACT = {"none": 0, "relu": 1, "swish": 2, "tanh": 3, "sigmoid": 4, "prelu": 5, "glu": 6, "leaky_relu": 7, "gelu": 8}

_SIGNATURES = {
    "aps_b200_abi_version": (c_int, []),
    "aps_b200_init": (c_int, [c_int]),
    "aps_b200_last_error": (c_int, [c_char_p, c_size_t]),
    "aps_b200_fft_table_floats": (c_int64, [c_int]),
    "aps_b200_fft_tables_host": (c_int, [c_int, c_int, c_void_p]),
    "aps_b200_num_frames": (c_int64, [c_int64, c_int, c_int, c_int]),
    "aps_b200_feats_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(StftDesc), POINTER(FeatDesc),
                                   c_void_p, c_void_p]),
    "aps_b200_stft_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(StftDesc), c_int, c_float,
                                  c_void_p, c_void_p]),
    "aps_b200_istft_num_samples": (c_int64, [c_int64, c_int, c_int, c_int]),
    "aps_b200_istft_fwd": (c_int, [c_void_p, c_int64, c_int64, POINTER(StftDesc), c_int, c_float, c_void_p,
                                   c_void_p]),
    "aps_b200_spec_feats_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_float,
                                        POINTER(FeatDesc), c_void_p, c_int64, c_void_p]),
    "aps_b200_ipd_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p, c_int64, c_int64, c_void_p]),
    "aps_b200_mask_colmax": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                     c_void_p, c_void_p]),
    "aps_b200_covar_fwd": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int64, c_int64, c_int64, c_int64,
                                   c_void_p, POINTER(c_int64), c_void_p, c_void_p, POINTER(c_int64), c_void_p,
                                   c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "aps_b200_mvdr_ref_logits": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int64, c_void_p, c_void_p]),
    "aps_b200_mvdr_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p,
                                      c_void_p]),
    "aps_b200_beamform_fwd": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int64, c_int64, c_int64, c_int64,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "aps_b200_linear_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, POINTER(Epilogue),
                                    c_void_p, c_int64, c_void_p]),
    "aps_b200_tf32_split": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "aps_b200_linear_tc_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                       POINTER(Epilogue), c_void_p, c_int64, c_void_p]),
    "aps_b200_linear_tc2_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                        POINTER(Epilogue), c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p]),
    "aps_b200_conv2d_nhwc_tc2_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                             c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(Epilogue),
                                             c_void_p, c_void_p, c_void_p]),
    "aps_b200_layernorm2_fwd": (c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_int64, c_float,
                                        c_void_p, c_void_p, c_float, c_int32, c_int64, c_int64, c_void_p, c_void_p,
                                        c_int64, c_void_p]),
    "aps_b200_dwconv1d2_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                       c_void_p, c_int, c_int, c_int, POINTER(Epilogue), c_void_p, c_void_p, c_int64,
                                       c_void_p, c_void_p]),
    "aps_b200_mhsa2_fwd": (c_int, [POINTER(AttnDesc), c_void_p, c_void_p, c_int64, c_void_p]),
    "aps_b200_conv2d_nhwc_tc_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                            c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(Epilogue),
                                            c_void_p, c_void_p]),
    "aps_b200_conv_transpose2d_nhwc_tc_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                                      c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                      POINTER(Epilogue), c_void_p, c_void_p]),
    "aps_b200_conv2d_nhwc_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, c_int, POINTER(Epilogue), c_void_p,
                                         c_void_p]),
    "aps_b200_conv_transpose2d_nhwc_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                                   c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                   POINTER(Epilogue), c_void_p, c_void_p]),
    "aps_b200_cmask_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_float, c_int,
                                   c_void_p, c_void_p]),
    "aps_b200_specaug_workspace_bytes": (c_int64, [c_int64]),
    "aps_b200_specaug_apply": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_int64,
                                       c_void_p, c_void_p]),
    "aps_b200_splice_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "aps_b200_delta_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                                   c_int64, c_int64, c_void_p]),
    "aps_b200_speed_perturb_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "aps_b200_layernorm_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_float,
                                       c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "aps_b200_dwconv1d_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                      c_void_p, c_int, c_int, c_int, POINTER(Epilogue), c_void_p, c_int64, c_void_p]),
    "aps_b200_mhsa_fwd": (c_int, [POINTER(AttnDesc), c_void_p, c_int64, c_void_p]),
    "aps_b200_cmvn_allband": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_void_p]),
    "aps_b200_utt_norm_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "aps_b200_utt_norm_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p,
                                      c_void_p, c_float, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "aps_b200_conv_transpose2d_nhwc_narrow_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                                          c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                          c_int, POINTER(Epilogue), c_void_p, c_void_p]),
    "aps_b200_lstm_group_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                        c_int64, c_int, c_void_p]),
    "aps_b200_lstm_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                  c_int64, c_void_p]),
    "aps_b200_lstm_group_tc_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                           c_int64, c_int, c_void_p, c_void_p]),
    "aps_b200_pair_objf_workspace_bytes": (c_int64, [c_int64, c_int64, c_int]),
    "aps_b200_pair_objf_fwd": (c_int, [POINTER(SignalList), POINTER(SignalList), c_int64, c_int64, POINTER(ObjfDesc),
                                       c_void_p, c_int64, c_void_p, c_void_p]),
}

_lib = None
_inited = set()


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m aps_b200.build` "
                               "(aps_b200 has no CPU / PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.aps_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("libaps_b200.so ABI version mismatch; rebuild with `python -m aps_b200.build`")
        _lib = lib
    return _lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().aps_b200_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


CALLS = 0          # C-ABI launch calls made so far (each is at least one kernel launch); read by bench.py


def check(rc: int) -> None:
    global CALLS
    CALLS += 1
    if rc != 0:
        raise RuntimeError(f"aps_b200: {last_error()} (code {rc})")


def require_cuda(t: th.Tensor, what: str = "input") -> th.device:
    """The product path is CUDA only; anything else is an error, never a fallback."""
    if not isinstance(t, th.Tensor) or not t.is_cuda:
        raise RuntimeError(f"aps_b200 kernels need a CUDA tensor for {what}; got "
                           f"{getattr(t, 'device', type(t))} (there is no CPU fallback)")
    dev = t.device
    if dev.index not in _inited:
        with th.cuda.device(dev):
            check(load().aps_b200_init(dev.index))
        _inited.add(dev.index)
    return dev


def stream_ptr(dev: th.device) -> int:
    return th.cuda.current_stream(dev).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def i64_array(values):
    """ctypes int64 array (e.g. tensor strides) for the C ABI."""
    return (c_int64 * len(values))(*[int(v) for v in values])


# ---------------------------------------------------------------------------- cached device tables
_tables: Dict[Tuple[int, int, int], th.Tensor] = {}


def fft_tables(nfft: int, inverse: bool, dev: th.device) -> th.Tensor:
    key = (nfft, int(inverse), dev.index)
    if key not in _tables:
        lib = load()
        n = lib.aps_b200_fft_table_floats(nfft)
        if n <= 0:
            raise RuntimeError(f"aps_b200: unsupported FFT size {nfft} (need a power of two in [64, 1024])")
        host = np.empty(n, dtype=np.float32)
        check(lib.aps_b200_fft_tables_host(nfft, int(inverse), host.ctypes.data))
        _tables[key] = th.from_numpy(host).to(dev)
    return _tables[key]
