"""Minimal (real, imag) pair with the part of the `aps.cplx.ComplexTensor` surface the hot path touches
(aps/cplx.py:18-198).  The kernels accept ANY object with `.real` / `.imag` tensors and return an
object of the SAME class as their complex input (so a reference `ComplexTensor` in gives a reference
`ComplexTensor` out); this class is only what is returned when the input was built from it."""
from numbers import Number
from typing import Optional

import torch as th


class ComplexTensor(object):

    def __init__(self, real: th.Tensor, imag: Optional[th.Tensor] = None, polar: bool = False) -> None:
        imag = th.zeros_like(real) if imag is None else imag
        if polar:
            real, imag = th.cos(imag) * real, th.sin(imag) * real
        self.real, self.imag = real, imag

    # -- structure ---------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.real.shape

    @property
    def device(self):
        return self.real.device

    @property
    def dtype(self):
        return self.real.dtype

    def size(self):
        return self.real.size()

    def dim(self) -> int:
        return self.real.dim()

    def _map(self, fn):
        return ComplexTensor(fn(self.real), fn(self.imag))

    def transpose(self, d0, d1):
        return self._map(lambda t: t.transpose(d0, d1))

    def view(self, *shape):
        return self._map(lambda t: t.view(*shape))

    def contiguous(self):
        return self._map(lambda t: t.contiguous())

    def to(self, *a, **k):
        return self._map(lambda t: t.to(*a, **k))

    def cpu(self):
        return self._map(lambda t: t.cpu())

    def cuda(self):
        return self._map(lambda t: t.cuda())

    def sum(self, dim=None, keepdim=False):
        return self._map(lambda t: t.sum(dim=dim, keepdim=keepdim))

    def __getitem__(self, item):
        return self._map(lambda t: t[item])

    def as_real(self) -> th.Tensor:
        return th.stack([self.real, self.imag], dim=-1)

    # -- arithmetic -----------------------------------------------------------------------------------
    def conj(self):
        return ComplexTensor(self.real, -1.0 * self.imag)

    def conj_transpose(self, d0, d1):
        return self.transpose(d0, d1).conj()

    def abs(self) -> th.Tensor:
        return (self.real**2 + self.imag**2).sqrt()

    def angle(self) -> th.Tensor:
        return th.atan2(self.imag, self.real)

    def __add__(self, o):
        if isinstance(o, (Number, th.Tensor)) and not isinstance(o, complex):
            return ComplexTensor(self.real + o, self.imag)
        return ComplexTensor(self.real + o.real, self.imag + o.imag)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, (Number, th.Tensor)) and not isinstance(o, complex):
            return ComplexTensor(self.real - o, self.imag)
        return ComplexTensor(self.real - o.real, self.imag - o.imag)

    def __mul__(self, o):
        if isinstance(o, (Number, th.Tensor)) and not isinstance(o, complex):
            return ComplexTensor(self.real * o, self.imag * o)
        return ComplexTensor(self.real * o.real - self.imag * o.imag, self.imag * o.real + self.real * o.imag)

    __rmul__ = __mul__


def is_complex_pair(x) -> bool:
    return not isinstance(x, th.Tensor) and hasattr(x, "real") and hasattr(x, "imag")


def like(x, real: th.Tensor, imag: th.Tensor):
    """Build a complex pair of the same class as `x` (falls back to this module's ComplexTensor)."""
    try:
        return type(x)(real, imag)
    except Exception:
        return ComplexTensor(real, imag)
