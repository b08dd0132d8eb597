"""Batched encoder pass for decoding — the batched form of `ASREncoderBase._batch_decoding_prep`
(/root/reference/aps/asr/ctc.py:58-84).

The reference loops over the utterances of a decoding batch (transform + encoder one by one) because zero padding changes
the result: padded frames leak into real ones through every convolution.  Here the whole ragged batch goes through the
fused feature kernel and the encoder ONCE: the transform's per-frame statistics never mix frames, the attention masks
padded keys, and `TransformerEncoder.forward_ragged` zeroes the frames beyond each utterance in front of every
convolution — so each utterance gets what it would get alone.  Chains whose result depends on the padded length (all-band
or global-statistics CMVN over padded frames, splice / delta across the boundary, pose "xl", utterance-level norms) fall
back to the reference's loop over the same modules.
"""
from typing import List, Optional, Tuple

import torch as th
import torch.nn as nn
from torch.nn.utils.rnn import pad_sequence


def _transform_is_framewise(asr_transform: Optional[nn.Module]) -> bool:
    """True when every output frame of the transform depends on its own samples only (and, for the last frames, on
    nothing beyond the utterance): the fused spectral chain with per-frame or global CMVN and nothing after it."""
    if asr_transform is None:
        return True
    from ..transform.asr import FeatureTransform, _match_tail
    if not isinstance(asr_transform, FeatureTransform) or asr_transform.spectra_index != 0:
        return False
    layers = list(asr_transform.transform)
    spec = layers[0]
    if getattr(spec, "center", False):
        return False                        # reflect padding at the END of the signal reads the padded samples
    tail = _match_tail(layers, 1)
    if tail is None or tail[5] != len(layers):
        return False
    cm = tail[3]
    if cm is not None and cm.gmean is None and not cm.per_band and (cm.norm_mean or cm.norm_var):
        return False                        # all-band statistics run over the padded frames too
    return True


def batch_decoding_prep(asr_transform: Optional[nn.Module], encoder: nn.Module, batch: List[th.Tensor],
                        batch_first: bool = True) -> Tuple[th.Tensor, th.Tensor]:
    """batch: list of waveforms S_i (or features T_i x F when `asr_transform` is None), all on one CUDA device.
    Returns (enc_out N x T x D (or T x N x D), enc_len N) exactly as aps/asr/ctc.py:58-84."""
    if not batch:
        raise RuntimeError("batch_decoding_prep: empty batch")
    exact = (_transform_is_framewise(asr_transform) and hasattr(encoder, "forward_ragged") and encoder.ragged_exact_ok()
             and all(b.dim() == batch[0].dim() for b in batch))
    if not exact:
        outs = []
        for inp in batch:                   # the reference's loop, on this package's modules
            if asr_transform is not None:
                inp, _ = asr_transform(inp[None, ...], None)
            else:
                inp = inp[None, ...]
            outs.append(encoder(inp, None)[0][0])
        enc_out = pad_sequence(outs, batch_first=False)
        enc_len = th.tensor([o.shape[0] for o in outs], device=enc_out.device)
        return (enc_out.transpose(0, 1) if batch_first else enc_out), enc_len
    lens = th.tensor([b.shape[0] for b in batch], dtype=th.int64)
    pad = pad_sequence(batch, batch_first=True)                         # N x S (or N x T x F), zero padded
    if asr_transform is not None:
        feats, nfr = asr_transform(pad, lens)
    else:
        feats, nfr = pad, lens
    enc_out, enc_len = encoder.forward_ragged(feats, nfr)
    T = int(enc_len.max())
    enc_out = enc_out[:, :T]
    enc_len = enc_len.to(enc_out.device)
    return (enc_out if batch_first else enc_out.transpose(0, 1)), enc_len
