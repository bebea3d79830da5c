"""spnn.Conv3d (TS/torchsparse/nn/modules/conv.py:15-72): same parameters (`kernel` (K,Cin,Cout) or (Cin,Cout) for a
1x1x1 kernel, optional `bias`), same initialisation, so reference checkpoints load unchanged."""
import math
from typing import List, Tuple, Union

import numpy as np
import torch
from torch import nn

from ...tensor import SparseTensor
from ...utils import make_ntuple
from .. import functional as F

__all__ = ['Conv3d']


class Conv3d(nn.Module):

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[int, List[int], Tuple[int, ...]] = 3,
                 stride: Union[int, List[int], Tuple[int, ...]] = 1, dilation: int = 1, bias: bool = False,
                 transposed: bool = False) -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = make_ntuple(kernel_size, ndim=3)
        self.stride = make_ntuple(stride, ndim=3)
        self.dilation = dilation
        self.transposed = transposed
        self.kernel_volume = int(np.prod(self.kernel_size))
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(*shape))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def extra_repr(self) -> str:
        parts = [f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}']
        if any(s != 1 for s in self.stride):
            parts.append(f'stride={self.stride}')
        if self.dilation != 1:
            parts.append(f'dilation={self.dilation}')
        if self.bias is None:
            parts.append('bias=False')
        if self.transposed:
            parts.append('transposed=True')
        return ', '.join(parts)

    def reset_parameters(self) -> None:
        fan = (self.out_channels if self.transposed else self.in_channels) * self.kernel_volume
        bound = 1 / math.sqrt(fan)
        self.kernel.data.uniform_(-bound, bound)
        if self.bias is not None:
            self.bias.data.uniform_(-bound, bound)

    def forward(self, input: SparseTensor) -> SparseTensor:
        return F.conv3d(input, self.kernel, kernel_size=self.kernel_size, bias=self.bias, stride=self.stride,
                        dilation=self.dilation, transposed=self.transposed)
