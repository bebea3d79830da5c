"""Building blocks of the MinkUNet family with the reference's module/parameter names
(R/pcseg/model/segmentor/voxel/minkunet/minkunet.py:23-180), so that reference checkpoints load unchanged:
  Basic(De)ConvolutionBlock: net = [Conv3d, BN, ReLU]
  ResidualBlock: net = [Conv3d, BN, ReLU, Conv3d, BN], downsample = Identity | [Conv3d 1x1, BN], relu
  Bottleneck:    net = [1x1, BN, 3^3, BN, 1x1(x4), BN] (not used by any TASeg config; module path only)
"""
from torch import nn

from .. import nn as spnn
from ..nn.modules.norm import fuse_bn_relu
from ..nn.utils import fapply
from ..tensor import SparseTensor


class SyncBatchNorm(nn.SyncBatchNorm):
    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)


class BatchNorm(spnn.BatchNorm):
    """spnn.BatchNorm: nn.BatchNorm1d on the feature rows (CUDA rows run the kernels of csrc/bn.cu)."""


def norm(channels: int, if_dist: bool) -> nn.Module:
    return SyncBatchNorm(channels) if if_dist else BatchNorm(channels)


class BasicConvolutionBlock(nn.Module):
    def __init__(self, inc, outc, ks=3, stride=1, dilation=1, if_dist=False):
        super().__init__()
        self.net = fuse_bn_relu(nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=ks, dilation=dilation, stride=stride),
                                              norm(outc, if_dist), spnn.ReLU(True)))

    def forward(self, x):
        return self.net(x)


class BasicDeconvolutionBlock(nn.Module):
    def __init__(self, inc, outc, ks=3, stride=1, if_dist=False):
        super().__init__()
        self.net = fuse_bn_relu(nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=ks, stride=stride, transposed=True),
                                              norm(outc, if_dist), spnn.ReLU(True)))

    def forward(self, x):
        return self.net(x)


class _Residual(nn.Module):
    expansion = 1

    def _shortcut(self, inc, outc, stride, if_dist):
        if inc == outc * self.expansion and stride == 1:
            return nn.Identity()
        return nn.Sequential(spnn.Conv3d(inc, outc * self.expansion, kernel_size=1, dilation=1, stride=stride),
                             norm(outc * self.expansion, if_dist))

    def forward(self, x):
        return self.relu(self.net(x) + self.downsample(x))


class ResidualBlock(_Residual):
    expansion = 1

    def __init__(self, inc, outc, ks=3, stride=1, dilation=1, if_dist=False):
        super().__init__()
        self.net = fuse_bn_relu(nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=ks, dilation=dilation, stride=stride),
                                              norm(outc, if_dist), spnn.ReLU(True),
                                              spnn.Conv3d(outc, outc, kernel_size=ks, dilation=dilation, stride=1),
                                              norm(outc, if_dist)))
        self.downsample = self._shortcut(inc, outc, stride, if_dist)
        self.relu = spnn.ReLU(True)


class Bottleneck(_Residual):
    expansion = 4

    def __init__(self, inc, outc, ks=3, stride=1, dilation=1, if_dist=False):
        super().__init__()
        self.net = nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=1, bias=False), norm(outc, if_dist),
                                 spnn.Conv3d(outc, outc, kernel_size=ks, stride=stride, bias=False, dilation=dilation),
                                 norm(outc, if_dist),
                                 spnn.Conv3d(outc, outc * self.expansion, kernel_size=1, bias=False),
                                 norm(outc * self.expansion, if_dist))
        self.downsample = self._shortcut(inc, outc, stride, if_dist)
        self.relu = spnn.ReLU(True)


BLOCKS = {'ResBlock': ResidualBlock, 'Bottleneck': Bottleneck}
