"""Test-only stand-in for librosa==0.8.1 (absent from this image).

Only `librosa.filters.mel` is needed to import the reference transform
(aps/transform/utils.py:10,148-154).  See oracle/ref_shims/README.md.
"""
from . import filters  # noqa: F401
