"""spnn.BatchNorm (TS/torchsparse/nn/modules/norm.py:10-13): BatchNorm1d over the (N, C) feature matrix."""
from torch import nn

from ...tensor import SparseTensor
from ..utils import fapply

__all__ = ['BatchNorm']


class BatchNorm(nn.BatchNorm1d):

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)
