"""spnn.BatchNorm (TS/torchsparse/nn/modules/norm.py:10-13): BatchNorm1d over the (N, C) feature matrix."""
from torch import nn

from ...tensor import SparseTensor
from ..utils import fapply

__all__ = ['BatchNorm', 'fuse_bn_relu']


class BatchNorm(nn.BatchNorm1d):

    fuse_relu = False     # set by fuse_bn_relu(): the ReLU module behind this BatchNorm is applied inside the same pass

    def _rows(self, feats):
        from ..functional.norm import batch_norm      # CUDA rows: streaming kernels of csrc/bn.cu; otherwise ATen
        out = batch_norm(self, feats, relu=self.fuse_relu)
        if out is not None:
            return out
        out = nn.BatchNorm1d.forward(self, feats)
        return nn.functional.relu(out) if self.fuse_relu else out

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, self._rows)


def fuse_bn_relu(seq: nn.Sequential) -> nn.Sequential:
    """Mark every (BatchNorm, ReLU) neighbour pair of `seq`: the BatchNorm applies the ReLU in its own pass (csrc/bn.cu,
    forward and backward) and the ReLU module becomes the identity.  Module and parameter names do not change."""
    from .activation import ReLU
    mods = list(seq)
    for a, b in zip(mods, mods[1:]):
        if isinstance(a, BatchNorm) and isinstance(b, ReLU) and not isinstance(a, nn.SyncBatchNorm):
            a.fuse_relu = True
            b.fused_upstream = True
    return seq
