"""Test-only stand-in for `soundfile` backed by scipy.io.wavfile (enough for
aps.io.read_audio on the reference's tests/data/transform/*.wav fixtures)."""
import numpy as np
from scipy.io import wavfile


def read(fname, start=0, stop=None, frames=-1, dtype="float64", always_2d=False, **kw):
    sr, data = wavfile.read(fname)
    if data.dtype == np.int16 and dtype in ("float32", "float64"):
        data = data.astype(dtype) / 32768.0
    else:
        data = data.astype(dtype)
    data = data[start:stop]
    if always_2d and data.ndim == 1:
        data = data[:, None]
    return data, sr


def write(fname, data, sr, **kw):
    wavfile.write(fname, sr, np.asarray(data))


class SoundFile:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("soundfile shim: SoundFile not supported")
