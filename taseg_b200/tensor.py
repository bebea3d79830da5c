"""SparseTensor / PointTensor containers with the reference's attribute surface (TS/torchsparse/tensor.py:10-105).

feats/F (N,C), coords/C (N,4) int32 [x,y,z,b], stride/s 3-tuple; `cmaps` {stride -> coords} and `kmaps`
{(stride, kernel, stride, dilation) -> kernel map} are shared by reference between all tensors derived from one
input, exactly like the reference, so pcseg code that copies them by hand keeps working.  Device-side coordinate
tables are cached inside `kmaps` under string-tagged keys so that they travel with it.
"""
from typing import Any, Dict, Tuple, Union

import torch

from .utils import make_ntuple

__all__ = ['SparseTensor', 'PointTensor']


class SparseTensor:

    def __init__(self, feats: torch.Tensor, coords: torch.Tensor, stride: Union[int, Tuple[int, ...]] = 1) -> None:
        self.feats = feats
        self.coords = coords
        self.stride = make_ntuple(stride, ndim=3)
        self.cmaps: Dict[Tuple[int, ...], torch.Tensor] = {}
        self.kmaps: Dict[Tuple[Any, ...], Any] = {}

    F = property(lambda self: self.feats, lambda self, v: setattr(self, 'feats', v))
    C = property(lambda self: self.coords, lambda self, v: setattr(self, 'coords', v))
    s = property(lambda self: self.stride, lambda self, v: setattr(self, 'stride', make_ntuple(v, ndim=3)))

    def _moved(self, fn):
        self.coords, self.feats = fn(self.coords), fn(self.feats)
        return self

    def cpu(self):
        return self._moved(lambda t: t.cpu())

    def cuda(self):
        return self._moved(lambda t: t.cuda())

    def detach(self):
        return self._moved(lambda t: t.detach())

    def to(self, device, non_blocking: bool = True):
        return self._moved(lambda t: t.to(device, non_blocking=non_blocking))

    def derive(self, feats, coords=None, stride=None) -> 'SparseTensor':
        """New tensor sharing this one's coordinate/kernel-map caches."""
        out = SparseTensor(feats, self.coords if coords is None else coords, self.stride if stride is None else stride)
        out.cmaps, out.kmaps = self.cmaps, self.kmaps
        return out

    def __add__(self, other):
        return self.derive(self.feats + other.feats)


class PointTensor:

    def __init__(self, feats, coords, idx_query=None, weights=None):
        self.F = feats
        self.C = coords
        self.idx_query = {} if idx_query is None else idx_query
        self.weights = {} if weights is None else weights
        self.additional_features = {'idx_query': {}, 'counts': {}}

    def cuda(self):
        self.F, self.C = self.F.cuda(), self.C.cuda()
        return self

    def detach(self):
        self.F, self.C = self.F.detach(), self.C.detach()
        return self

    def to(self, device, non_blocking=True):
        self.F, self.C = self.F.to(device, non_blocking=non_blocking), self.C.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        out = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        out.additional_features = self.additional_features
        return out
