"""spnn.ReLU / LeakyReLU (TS/torchsparse/nn/modules/activation.py:9-18)."""
from torch import nn

from ...tensor import SparseTensor
from ..utils import fapply

__all__ = ['ReLU', 'LeakyReLU']


class ReLU(nn.ReLU):
    fused_upstream = False     # True: the BatchNorm in front of this module already applied the ReLU (norm.fuse_bn_relu)

    def forward(self, input: SparseTensor) -> SparseTensor:
        if self.fused_upstream:
            return input
        return fapply(input, super().forward)


class LeakyReLU(nn.LeakyReLU):

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)
