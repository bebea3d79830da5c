"""spnn.BatchNorm (TS/torchsparse/nn/modules/norm.py:10-13): BatchNorm1d over the (N, C) feature matrix."""
from torch import nn

from ...tensor import SparseTensor
from ..utils import fapply

__all__ = ['BatchNorm']


class BatchNorm(nn.BatchNorm1d):

    def _rows(self, feats):
        from ..functional.norm import batch_norm      # CUDA rows: streaming kernels of csrc/bn.cu; otherwise ATen
        out = batch_norm(self, feats)
        return out if out is not None else nn.BatchNorm1d.forward(self, feats)

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, self._rows)
