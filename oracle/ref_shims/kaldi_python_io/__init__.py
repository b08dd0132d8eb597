"""Test-only stand-in: the reference imports these names (aps/transform/asr.py:28,
aps/loader/*) but the hot path never calls them."""


def _absent(*a, **k):
    raise RuntimeError("kaldi_python_io is not installed in this image (shim)")


class Reader:
    def __init__(self, *a, **k):
        _absent()


ScriptReader = ArchiveReader = ArchiveWriter = Nnet3EgsReader = AlignArchiveReader = Reader
