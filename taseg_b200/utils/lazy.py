"""A device scalar that becomes a Python float when it is first read.

The reference's model forward returns `loss.item()` twice in its display dictionaries
(R/pcseg/model/segmentor/voxel/minkunet/minkunet_ms.py:429-431): a host synchronisation between forward and backward of
every training step, during which nothing is queued for the GPU.  LazyScalar keeps the dictionaries' contract — formatting,
arithmetic, comparison and float() all work — but the wait happens when (and if) the logger reads the value."""
import torch

__all__ = ['LazyScalar']


class LazyScalar:
    __slots__ = ('_t', '_v')

    def __init__(self, t: torch.Tensor):
        self._t, self._v = t.detach(), None

    def __float__(self) -> float:
        if self._v is None:
            self._v, self._t = float(self._t), None
        return self._v

    def item(self) -> float:
        return float(self)

    def __int__(self):
        return int(float(self))

    def __bool__(self):
        return bool(float(self))

    def __round__(self, n=None):
        return round(float(self), n)

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return repr(float(self))

    __str__ = __repr__

    def __hash__(self):
        return hash(float(self))

    def __array__(self, dtype=None, copy=None):
        import numpy as np
        return np.asarray(float(self), dtype=dtype)

    def __neg__(self):
        return -float(self)

    def __abs__(self):
        return abs(float(self))


def _binary(name):
    def op(self, other):
        return getattr(float(self), name)(float(other) if isinstance(other, LazyScalar) else other)
    op.__name__ = name
    return op


for _n in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__', '__truediv__', '__rtruediv__', '__pow__',
           '__rpow__', '__lt__', '__le__', '__gt__', '__ge__', '__eq__', '__ne__'):
    setattr(LazyScalar, _n, _binary(_n))
