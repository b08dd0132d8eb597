"""aps_b200 — B200 (sm_100a) implementation of the APS data-parallel speech hot path.

Python shells that keep the reference's nn.Module surfaces (aps.transform.AsrTransform /
EnhTransform, ...) over hand-written CUDA kernels reached through the C ABI in
`include/aps_b200.h` (libaps_b200.so, built in-tree by `python -m aps_b200.build`).
There is no CPU or PyTorch fallback: CUDA tensors and the built library are required.
"""
__version__ = "0.1.0"


def register_into_aps() -> None:
    """Take over the reference's registry entries so unmodified recipes pick up this package.

    `aps.libs.Register.register` overwrites on a duplicate alias (aps/libs.py:26-35), so calling
    this once after `import aps` makes `aps_transform("asr" | "enh")` (aps/libs.py:150-155) return
    the classes of this package.  See INTEGRATION.md.
    """
    import warnings

    from aps.libs import ApsModules, ApsRegisters  # the reference, must be importable

    from .transform import AsrTransform, EnhTransform
    ApsModules.transform.import_all()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ApsRegisters.transform.register("asr")(AsrTransform)
        ApsRegisters.transform.register("enh")(EnhTransform)
