"""CPU oracle for the backbone forwards (eval mode): MinkUNet, MinkUNetMs, SPVCNN.

TEST INFRASTRUCTURE ONLY (see oracle/ts_oracle.py header).  R/ = /root/reference/.

Each network is restated as a walk over the reference's state_dict names, so the same
checkpoint feeds the reference, this oracle and the CUDA product.  Pinned against the real
reference models by tests/golden/make_golden.py (logits fixtures).

`ops` selects who executes the native entry points: the numpy restatement (ts_oracle) or the
reference's compiled CPU backend (oracle/ref_backend.RefOps) — the latter is what
`bench.py --impl reference` times.
"""
from __future__ import annotations

import numpy as np

from . import ts_oracle as T

F32 = np.float32


class NumpyOps:
    conv3d = staticmethod(T.conv3d)
    initial_voxelize = staticmethod(T.initial_voxelize)
    point_to_voxel = staticmethod(T.point_to_voxel)
    voxel_to_point = staticmethod(T.voxel_to_point)

    @staticmethod
    def matmul(a, b):
        return np.asarray(a, F32) @ np.asarray(b, F32)


class Net:
    """Eval-mode walker.  sd: name -> np.ndarray (float32)."""

    def __init__(self, sd: dict, ops=NumpyOps):
        self.sd = {k: np.asarray(v) for k, v in sd.items()}
        self.ops = ops

    # -- layer primitives ---------------------------------------------------------------
    def bn(self, f, p, eps=1e-5):
        """nn.BatchNorm1d eval (R/.../minkunet.py:27-29 via fapply)."""
        sd = self.sd
        inv = F32(1.0) / np.sqrt(sd[p + ".running_var"].astype(F32) + F32(eps))
        return ((f - sd[p + ".running_mean"]) * inv * sd[p + ".weight"] + sd[p + ".bias"]).astype(F32)

    def conv(self, x, p, ks, stride=1, transposed=False):
        return self.ops.conv3d(x, self.sd[p + ".kernel"], ks, stride=stride, transposed=transposed)

    def conv_bn_relu(self, x, p, ks, stride=1, transposed=False, relu=True):
        """BasicConvolutionBlock / BasicDeconvolutionBlock (minkunet.py:31-80): p.0 conv, p.1 BN, ReLU."""
        y = self.conv(x, p + ".0", ks, stride, transposed)
        y.F = self.bn(y.F, p + ".1")
        if relu:
            y.F = np.maximum(y.F, 0)
        return y

    def resblock(self, x, p):
        """ResidualBlock (minkunet.py:83-129): net = conv3-BN-ReLU-conv3-BN; 1x1 conv+BN shortcut when
        the state_dict has p.downsample.0.kernel; relu(net(x)+shortcut(x))."""
        y = self.conv(x, p + ".net.0", 3)
        y.F = np.maximum(self.bn(y.F, p + ".net.1"), 0)
        y = self.conv(y, p + ".net.3", 3)
        y.F = self.bn(y.F, p + ".net.4")
        if p + ".downsample.0.kernel" in self.sd:
            s = x.like(self.ops.matmul(x.F, self.sd[p + ".downsample.0.kernel"]))
            s.F = self.bn(s.F, p + ".downsample.1")
        else:
            s = x
        return y.like(np.maximum(y.F + s.F, 0))

    def _nblocks(self, prefix, start):
        n = 0
        while f"{prefix}.{start + n}.net.0.kernel" in self.sd:
            n += 1
        return n

    def stage(self, x, p):
        """stageN = [BasicConvolutionBlock(ks2,s2)] + ResidualBlocks (minkunet_ms.py:226-273)."""
        x = self.conv_bn_relu(x, p + ".0.net", 2, stride=2)
        for i in range(self._nblocks(p, 1)):
            x = self.resblock(x, f"{p}.{1 + i}")
        return x

    def up(self, x, skip, p):
        """upN = [BasicDeconvolutionBlock(ks2,s2), Sequential(ResidualBlocks)] with a channel concat of
        the encoder skip in between (minkunet_ms.py:275-333,400-417)."""
        y = self.conv_bn_relu(x, p + ".0.net", 2, stride=2, transposed=True)
        y = y.like(np.concatenate([y.F, skip.F], 1))
        for i in range(self._nblocks(p + ".1", 0)):
            y = self.resblock(y, f"{p}.1.{i}")
        return y

    def stem(self, x):
        x = self.conv(x, "stem.0", 3)
        x.F = np.maximum(self.bn(x.F, "stem.1"), 0)
        x = self.conv(x, "stem.3", 3)
        x.F = np.maximum(self.bn(x.F, "stem.4"), 0)
        return x

    def classifier(self, f):
        return (self.ops.matmul(f, self.sd["classifier.0.weight"].T) + self.sd["classifier.0.bias"]).astype(F32)

    def point_mlp(self, f, i):
        """SPVCNN point_transforms[i] = Linear-BN-ReLU (spvcnn.py:335-351)."""
        p = f"point_transforms.{i}"
        y = self.ops.matmul(f, self.sd[p + ".0.weight"].T) + self.sd[p + ".0.bias"]
        return np.maximum(self.bn(y.astype(F32), p + ".1"), 0)

    # -- networks -----------------------------------------------------------------------
    def _unet(self, x0, z_first, spv=False):
        ops = self.ops
        z0 = ops.voxel_to_point(x0, z_first, nearest=False)
        x1 = ops.point_to_voxel(x0, z0) if spv else x0
        x1 = self.stage(x1, "stage1")
        x2 = self.stage(x1, "stage2")
        x3 = self.stage(x2, "stage3")
        x4 = self.stage(x3, "stage4")
        z1 = ops.voxel_to_point(x4, z0)
        if spv:
            z1.F = z1.F + self.point_mlp(z0.F, 0)
            y1 = ops.point_to_voxel(x4, z1)
        else:
            y1 = x4
        y1 = self.up(y1, x3, "up1")
        y2 = self.up(y1, x2, "up2")
        z2 = ops.voxel_to_point(y2, z1)
        if spv:
            z2.F = z2.F + self.point_mlp(z1.F, 1)
            y3 = ops.point_to_voxel(y2, z2)
        else:
            y3 = y2
        y3 = self.up(y3, x1, "up3")
        y4 = self.up(y3, x0, "up4")
        z3 = ops.voxel_to_point(y4, z2)
        if spv:
            z3.F = z3.F + self.point_mlp(z2.F, 2)
        return self.classifier(np.concatenate([z1.F, z2.F, z3.F], 1))

    def minkunet_ms(self, coords, feats):
        """R/pcseg/model/segmentor/voxel/minkunet/minkunet_ms.py:385-420 (dropout p=0 in eval).
        coords (N,4) int32 deduped voxels [x,y,z,b]; feats (N,Cin) -> logits per voxel."""
        x = T.SparseTensor(np.asarray(feats, F32), np.asarray(coords, np.int32))
        z = T.PointTensor(x.F, x.C.astype(F32))
        return self._unet(self.stem(x), z)

    def minkunet(self, coords, feats, pres=0.05, vres=0.05):
        """R/pcseg/model/segmentor/voxel/minkunet/minkunet.py:385-422."""
        z = T.PointTensor(np.asarray(feats, F32), np.asarray(coords, np.int32).astype(F32))
        x0 = self.ops.initial_voxelize(z, pres, vres)
        return self._unet(self.stem(x0), z)

    def spvcnn(self, coords, feats, pres=0.05, vres=0.05):
        """R/pcseg/model/segmentor/fusion/spvcnn/spvcnn.py:399-449."""
        z = T.PointTensor(np.asarray(feats, F32), np.asarray(coords, np.int32).astype(F32))
        x0 = self.ops.initial_voxelize(z, pres, vres)
        return self._unet(self.stem(x0), z, spv=True)


def gather_points(logits, inverse_map, point_mask=None, num_points=None):
    """Eval tail, one sample: out[inverse_map][point_mask][:num_points]
    (minkunet_ms.py:441-456 / minkunet.py:436-455)."""
    o = logits[inverse_map]
    if point_mask is not None:
        o = o[point_mask]
    return o if num_points is None else o[:num_points]
