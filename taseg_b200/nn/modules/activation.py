"""spnn.ReLU / LeakyReLU (TS/torchsparse/nn/modules/activation.py:9-18)."""
from torch import nn

from ...tensor import SparseTensor
from ..utils import fapply

__all__ = ['ReLU', 'LeakyReLU']


class ReLU(nn.ReLU):

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)


class LeakyReLU(nn.LeakyReLU):

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)
