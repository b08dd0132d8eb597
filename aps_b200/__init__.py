"""aps_b200 — B200 (sm_100a) implementation of the APS data-parallel speech hot path.

Python shells that keep the reference's nn.Module surfaces (aps.transform.AsrTransform /
EnhTransform, ...) over hand-written CUDA kernels reached through the C ABI in
`include/aps_b200.h` (libaps_b200.so, built in-tree by `python -m aps_b200.build`).
There is no CPU or PyTorch fallback: CUDA tensors and the built library are required.
"""
__version__ = "0.1.0"


def register_into_aps(verbose: bool = False) -> dict:
    """Take over the reference's registry entries and class bindings so UNMODIFIED recipes pick up this package.

    `aps.libs.Register.register` overwrites on a duplicate alias (aps/libs.py:26-35), so calling this once after
    `import aps` makes `aps_transform("asr" | "enh")`, `aps_sse_nnet("sse@dccrn" | "sse@freq_tcn" | "sse@time_tcn" |
    "sse@freq_xfmr")` and `aps_task("sse@sisnr" | "sse@snr" | "sse@wa" | "sse@freq_linear_sa" | ...)` (aps/libs.py:124-173)
    return the classes of this package.  Networks that build their encoder / beamformer by NAME keep their own class and
    get ours underneath: `aps.asr.ctc.CtcASR` (`asr@ctc`, aps/asr/ctc.py:15, :48) and `aps.sse.bss.transformer.FreqXfmr`
    look `TransformerEncoder` up in their module globals, `aps.asr.filter.mvdr.RNNMaskMvdr` (`rnn_mask_mvdr`) looks up
    `MvdrBeamformer` — those globals are re-bound.  Returns {what: replacement} for logging.  See INTEGRATION.md."""
    import importlib
    import warnings

    from aps.libs import ApsModules, ApsRegisters  # the reference, must be importable

    from .asr.filter import MvdrBeamformer
    from .asr.transformer import TransformerEncoder
    from .sse.bss import DCCRN, FreqConvTasNet, FreqXfmr, TimeConvTasNet
    from . import task as T
    from .transform import AsrTransform, EnhTransform
    done = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ApsModules.transform.import_all()
        for alias, cls in (("asr", AsrTransform), ("enh", EnhTransform)):
            ApsRegisters.transform.register(alias)(cls)
            done[f"transform:{alias}"] = cls
        ApsModules.sse.import_all()
        for alias, cls in (("sse@dccrn", DCCRN), ("sse@freq_tcn", FreqConvTasNet), ("sse@time_tcn", TimeConvTasNet),
                           ("sse@freq_xfmr", FreqXfmr)):
            ApsRegisters.sse.register(alias)(cls)
            done[f"sse:{alias}"] = cls
        ApsModules.task.import_all()
        for alias, cls in (("sse@sisnr", T.SisnrTask), ("sse@snr", T.SnrTask), ("sse@wa", T.WaTask),
                           ("sse@freq_linear_sa", T.LinearFreqSaTask), ("sse@freq_mel_sa", T.MelFreqSaTask),
                           ("sse@time_linear_sa", T.LinearTimeSaTask), ("sse@time_mel_sa", T.MelTimeSaTask),
                           ("sse@complex_mapping", T.ComplexMappingTask), ("sse@complex_masking", T.ComplexMaskingTask)):
            ApsRegisters.task.register(alias)(cls)
            done[f"task:{alias}"] = cls
        ApsModules.asr.import_all()
        for mod, attr, cls in (("aps.asr.transformer.encoder", "TransformerEncoder", TransformerEncoder),
                               ("aps.asr.ctc", "TransformerEncoder", TransformerEncoder),
                               ("aps.sse.bss.transformer", "TransformerEncoder", TransformerEncoder),
                               ("aps.sse.bss.sepformer", "TransformerEncoder", TransformerEncoder),
                               ("aps.asr.filter.mvdr", "MvdrBeamformer", MvdrBeamformer)):
            m = importlib.import_module(mod)
            if not hasattr(m, "_aps_b200_saved_" + attr):
                setattr(m, "_aps_b200_saved_" + attr, getattr(m, attr))
            setattr(m, attr, cls)
            done[f"{mod}.{attr}"] = cls
    if verbose:
        for k, v in done.items():
            print(f"aps_b200: {k} -> {v.__module__}.{v.__name__}")
    return done


def unregister_from_aps() -> None:
    """Undo the module-global re-bindings of `register_into_aps` and re-import the reference's own registry entries
    (used by the tests so the live reference stays available as the oracle)."""
    import importlib
    import sys
    import warnings
    for mod in list(sys.modules.values()):
        if mod is None or not getattr(mod, "__name__", "").startswith("aps."):
            continue
        for attr in ("TransformerEncoder", "MvdrBeamformer"):
            saved = getattr(mod, "_aps_b200_saved_" + attr, None)
            if saved is not None:
                setattr(mod, attr, saved)
                delattr(mod, "_aps_b200_saved_" + attr)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name in ("aps.transform.asr", "aps.transform.enh", "aps.sse.bss.dccrn", "aps.sse.bss.tcn", "aps.sse.bss.transformer",
                     "aps.task.sse"):
            if name in sys.modules:
                importlib.reload(sys.modules[name])
