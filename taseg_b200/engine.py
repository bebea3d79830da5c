"""Eval-mode inference engine: the backbone forward as a fused bf16 tcgen05 pipeline.

`Engine(model)` walks a MinkUNet / MinkUNetMs / SPVCNN from taseg_b200.segmentor (reference parameter names) and
compiles every Conv3d+BatchNorm(+ReLU)(+residual add) group into ONE tensor-core launch:
  * BatchNorm running statistics are folded: scale into the packed bf16 weights, shift into the epilogue bias;
  * ReLU and the residual add run in the epilogue on the fp32 accumulator read from TMEM;
  * `torchsparse.cat([y, skip])` is never materialised — the next convolution gathers from both tensors;
  * 1x1x1 shortcut convolutions and point MLPs use the same kernel with the identity map;
  * the classifier is applied per scale at VOXEL level (Linear is linear: devox(F) @ W == devox(F @ W)), so the
    trilinear devoxelisation moves num_class instead of 256+128+96 channels per point (voxel nets only).
Activations stay bf16 (N, C) with C padded to a multiple of 16; accumulation is fp32.  Results agree with the fp32
module path to bf16 round-off (tests state the bound); kernel maps / voxel sets are the same bit-exact objects.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from torch import nn

from . import ops
from .nn.modules.conv import Conv3d
from .nn.utils.kernel import kernel_offsets_np
from .segmentor.unet import _SparseUNet

pad16 = ops.pad16


def _fold(bn: Optional[nn.Module], c_out: int, device):
    if bn is None:
        return torch.ones(c_out, device=device), torch.zeros(c_out, device=device)
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return scale, bn.bias.detach().float() - bn.running_mean.detach().float() * scale


SORT_ROWS = os.environ.get("TSG_SORT_ROWS", "1") != "0"   # mask-sorted tile rows (A/B switch for profiling)


def conv_map(km, transposed: bool = False):
    """(nbr, tile_mask, perm, split) the tensor-core convolution consumes for kernel map `km` (split: K-split work items
    of launches with about as many tiles as SMs, or None)."""
    if SORT_ROWS:
        return km.sorted(transposed) + (km.split_items(transposed),)
    return (km.nbr_t if transposed else km.nbr, km.tile_mask(transposed), None, None)


FOLD_SHORTCUT = os.environ.get("TSG_FOLD_SHORTCUT", "1") != "0"   # A/B switch: 1x1 shortcut as a second K phase of conv 2


def _fold_pack(kernel: torch.Tensor, bn, r0: int, r1: int, c_out_pad: int, bias: Optional[torch.Tensor] = None):
    """(K, r0 + r1, c_out) kernel + BN -> (packed bf16 image with the BN scale folded in, fp32 shift (c_out_pad))."""
    w = kernel.detach().float()
    if w.dim() == 2:
        w = w.unsqueeze(0)
    k, c_in, c_out = w.shape
    assert c_in == r0 + r1, (c_in, r0, r1)
    c0, c1 = pad16(r0), pad16(r1) if r1 else 0
    scale, shift = _fold(bn, c_out, w.device)
    if bias is not None:
        shift = shift + bias.detach().float() * scale
    wp = torch.zeros((k, c0 + c1, c_out_pad), device=w.device)
    wp[:, :r0, :c_out] = w[:, :r0] * scale
    if r1:
        wp[:, c0:c0 + r1, :c_out] = w[:, r0:] * scale
    sh = torch.zeros(c_out_pad, device=w.device)
    sh[:c_out] = shift
    return ops.pack_weights(wp, c0, c1), sh, k, c0, c1


class FusedConv:
    """Conv3d (+BN) compiled for tsg_conv_fwd_tc2.  r0/r1 = real channels of the (up to) two input tensors.
    shortcut = (kernel, bn, s0, s1): a 1x1x1 convolution (+BN) of a second pair of tensors with s0/s1 real channels,
    accumulated into the same output tile (ResidualBlock's downsample branch)."""

    def __init__(self, kernel: torch.Tensor, bn, r0: int, r1: int = 0, relu: bool = False, bias: Optional[torch.Tensor] = None,
                 shortcut=None):
        c_out = kernel.shape[-1]
        self.c_out, self.relu, self.c_out_pad = c_out, relu, pad16(c_out)
        self.packed, self.bias, self.k, self.c0, self.c1 = _fold_pack(kernel, bn, r0, r1, self.c_out_pad, bias)
        self.sc_packed = None
        if shortcut is not None:
            sk, sbn, s0, s1 = shortcut
            assert sk.shape[-1] == c_out
            self.sc_packed, sc_shift, _, self.sc0, self.sc1 = _fold_pack(sk, sbn, s0, s1, self.c_out_pad)
            self.bias = self.bias + sc_shift

    def __call__(self, x0, x1, m, n_out, residual=None, out_dtype=torch.bfloat16, sc_in=None, n_dev=None):
        """m = (nbr, tile_mask, perm) from conv_map(), or None for the identity map (1x1x1 convolutions, point MLPs).
        sc_in = (s0, s1 | None): the inputs of the folded shortcut.  n_dev: device row counter (n_out is then a capacity)."""
        nbr, tile_mask, perm = m[:3] if m is not None else (None, None, None)
        split = m[3] if m is not None and len(m) > 3 else None
        shortcut = None
        if self.sc_packed is not None:
            centre = nbr[self.k // 2] if perm is not None else None     # identity in tile-row order = the centre offset's line
            shortcut = (sc_in[0], sc_in[1], self.sc_packed, centre)
        return ops.conv_forward_tc(x0, x1, self.packed, self.k, self.c_out_pad, nbr, tile_mask, n_out, bias=self.bias,
                                   residual=residual, relu=self.relu, out_dtype=out_dtype, perm=perm, shortcut=shortcut,
                                   n_dev=n_dev, split=split)


class FusedBlock:
    """ResidualBlock: conv-BN-ReLU, conv-BN, (+ 1x1 conv-BN shortcut), add, ReLU."""

    def __init__(self, block, r0: int, r1: int = 0):
        net = block.net
        assert isinstance(net[0], Conv3d) and len(net) == 5, "engine supports ResBlock (the only block TASeg configures)"
        self.a = FusedConv(net[0].kernel, net[1], r0, r1, relu=True)
        self.shortcut, self.folded = None, False
        has_sc = not isinstance(block.downsample, nn.Identity)
        fold = has_sc and FOLD_SHORTCUT and net[3].kernel.dim() == 3 and net[3].kernel.shape[0] == 27
        if fold:    # shortcut as extra K slices of the second convolution (same accumulator, one launch less)
            self.b = FusedConv(net[3].kernel, net[4], net[3].in_channels, 0, relu=True,
                               shortcut=(block.downsample[0].kernel, block.downsample[1], r0, r1))
            self.folded = True
        else:
            self.b = FusedConv(net[3].kernel, net[4], net[3].in_channels, 0, relu=True)
            if has_sc:
                self.shortcut = FusedConv(block.downsample[0].kernel, block.downsample[1], r0, r1, relu=False)

    def __call__(self, x0, x1, m, n, n_dev=None):
        """m = (nbr, tile_mask, perm) of the level's 3x3x3 map, n rows (capacity when n_dev is given)."""
        h = self.a(x0, x1, m, n, n_dev=n_dev)
        if self.folded:
            return self.b(h, None, m, n, sc_in=(x0, x1), n_dev=n_dev)
        res = self.shortcut(x0, x1, None, n, n_dev=n_dev) if self.shortcut is not None else x0
        return self.b(h, None, m, n, residual=res, n_dev=n_dev)


class Level:
    """One pyramid level: n rows (exact, or the buffer capacity when the count lives on the device in n_dev)."""
    __slots__ = ("stride", "coords", "n", "n_dev", "table", "km3", "km2", "m3", "m2", "m2t")

    def map3(self):
        """(nbr, tile_mask, perm) of the 3x3x3 stride-1 map at this level."""
        return self.m3 if self.m3 is not None else conv_map(self.km3)

    def map2(self):
        """... of the 2x2x2 stride-2 map to the next coarser level (rows = coarse voxels)."""
        return self.m2 if self.m2 is not None else conv_map(self.km2)

    def map2t(self):
        """... of its transpose (rows = this level's voxels): the transposed convolution of the decoder."""
        return self.m2t if self.m2t is not None else conv_map(self.km2, True)


class Geometry:
    """Voxel pyramid of one batch: coordinates, tables and kernel maps at tensor strides 1..16."""

    def __init__(self, coords: torch.Tensor, n_levels: int = 5, field_bits=None):
        self.levels: List[Level] = []
        c = coords.contiguous()
        for l in range(n_levels):
            lv = Level()
            lv.n_dev = lv.m3 = lv.m2 = lv.m2t = None
            lv.stride, lv.coords, lv.n = 2 ** l, c, c.shape[0]
            lv.table = ops.Table.from_coords(c)
            lv.km3 = ops.build_kmap(lv.table, lv.n, c, kernel_offsets_np(3, lv.stride))
            lv.km2 = None
            if l + 1 < n_levels:
                nxt = ops.unique_coords(c, trunc_stride=2 * lv.stride, field_bits=field_bits)
                lv.km2 = ops.build_kmap(lv.table, lv.n, nxt, kernel_offsets_np(2, lv.stride))
                c = nxt
            self.levels.append(lv)


class Engine:
    def __init__(self, model: _SparseUNet):
        assert not model.training, "Engine folds BatchNorm running statistics: call model.eval() first"
        self.model = model
        self.spv, self.voxelize_input = model.point_branch, model.voxelize_input
        self.in_dim, self.num_class = model.in_feature_dim, model.num_class
        self.pres, self.vres = model.pres, model.vres
        st = model.stem
        c = st[0].out_channels
        self.stem = [FusedConv(st[0].kernel, st[1], self.in_dim, relu=True), FusedConv(st[3].kernel, st[4], c, relu=True)]
        self.down, self.enc = [], []
        width = c
        skip_w = [c]
        for i in range(4):
            stage = getattr(model, f"stage{i + 1}")
            self.down.append(FusedConv(stage[0].net[0].kernel, stage[0].net[1], width, relu=True))
            blocks = []
            for blk in list(stage)[1:]:
                blocks.append(FusedBlock(blk, width))
                width = blk.net[3].out_channels
            self.enc.append(blocks)
            skip_w.append(width)
        self.up, self.dec = [], []
        for i in range(4):
            up = getattr(model, f"up{i + 1}")
            self.up.append(FusedConv(up[0].net[0].kernel, up[0].net[1], width, relu=True))
            width = up[0].net[0].out_channels
            skip = skip_w[3 - i]
            blocks = []
            for j, blk in enumerate(up[1]):
                blocks.append(FusedBlock(blk, width, skip) if j == 0 else FusedBlock(blk, width))
                width = blk.net[3].out_channels
            self.dec.append(blocks)
        lin = model.classifier[0]
        self.head_dims = [skip_w[4], self.dec[1][-1].b.c_out, self.dec[3][-1].b.c_out]
        wt = lin.weight.detach().float().t().contiguous()                 # (480, num_class)
        if self.spv:
            self.head = FusedConv(wt, None, wt.shape[0], bias=lin.bias)
            self.mlps = []
            for seq in model.point_transforms:
                self.mlps.append(FusedConv(seq[0].weight.detach().float().t().contiguous(), seq[1], seq[0].in_features,
                                           relu=True, bias=seq[0].bias))
        else:
            offs = [0, self.head_dims[0], self.head_dims[0] + self.head_dims[1]]
            self.heads = [FusedConv(wt[offs[i]:offs[i] + self.head_dims[i]], None, self.head_dims[i],
                                    bias=lin.bias if i == 2 else None) for i in range(3)]

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, coords: torch.Tensor, feats: torch.Tensor, return_geometry: bool = False, field_bits=None,
                 out_rows: Optional[torch.Tensor] = None):
        """coords (N,4) int32 [x,y,z,b], feats (N,>=in_dim) fp32 -> logits (N, num_class) fp32, one row per input row.
        field_bits: optional (bx,by,bz,bb) promise 0 <= coordinate < 2^bits (shorter radix sorts in the pyramid).
        out_rows: optional (M,) int32 — return logits[out_rows] only (the eval gather of minkunet_ms.py:441-456 fused
        into the last kernel)."""
        feats = feats[:, :self.in_dim].float().contiguous()
        zc = coords.float().contiguous()
        if self.voxelize_input:      # initial_voxelize (minkunet/utils.py:11-36)
            zc, floor_c = ops.rescale_coords(zc, self.pres, self.vres)
            vox, _, inv0 = ops.unique_coords(floor_c, want_index=True, want_inverse=True, by_hash=True)
            cnt0 = ops.spcount(inv0, vox.shape[0])
            x_f = ops.voxelize_forward(feats, inv0, cnt0)
            vox = vox.contiguous()
        else:
            vox, x_f = coords.contiguous(), feats
        geo = Geometry(vox, field_bits=field_bits if not self.voxelize_input else None)
        aux = dict(zc=zc, inv0=inv0, cnt0=cnt0) if self.voxelize_input else dict(zc=zc)
        logits = self.run(geo, x_f, aux, out_rows)
        return (logits, geo) if return_geometry else logits

    @torch.no_grad()
    def run(self, geo, x_f: torch.Tensor, aux: dict, out_rows: Optional[torch.Tensor] = None):
        """The backbone over a built pyramid `geo` (Geometry, or pipeline.GeometryDev whose row counts stay on the device)."""
        L = geo.levels
        zc = aux["zc"]
        nd = [lv.n_dev for lv in L]
        x = ops.cast_pad_bf16(x_f, self.stem[0].c0, n_dev=nd[0])
        for conv in self.stem:
            x = conv(x, None, L[0].map3(), L[0].n, n_dev=nd[0])
        x0 = x
        if self.spv:
            inv0, cnt0 = aux["inv0"], aux["cnt0"]
            q1 = ops.trilinear_query(L[0].table, zc, 1)
            q16 = ops.trilinear_query(L[4].table, zc, 16)
            q4 = ops.trilinear_query(L[2].table, zc, 4)
            z0 = ops.devoxelize_forward(x0, *q1)
            p2v1 = (inv0, cnt0)
            x = ops.voxelize_forward(z0, *p2v1)
        skips = [x0]
        for i in range(4):
            x = self.down[i](x, None, L[i].map2(), L[i + 1].n, n_dev=nd[i + 1])
            m3 = L[i + 1].map3()
            for blk in self.enc[i]:
                x = blk(x, None, m3, L[i + 1].n, nd[i + 1])
            skips.append(x)
        x4 = x
        if self.spv:
            z1 = ops.devoxelize_forward(x4, *q16) + self.mlps[0](z0, None, None, z0.shape[0])
            i16 = ops.point_query(L[4].table, zc, 16)
            x = ops.voxelize_forward(z1, i16, ops.spcount(i16, L[4].n))
        ys = []
        for i in range(4):
            lv = L[3 - i]
            x = self.up[i](x, None, lv.map2t(), lv.n, n_dev=lv.n_dev)
            skip = skips[3 - i]
            m3 = lv.map3()
            for j, blk in enumerate(self.dec[i]):
                x = blk(x, skip if j == 0 else None, m3, lv.n, lv.n_dev)
            ys.append(x)
            if self.spv and i == 1:
                z2 = ops.devoxelize_forward(x, *q4) + self.mlps[1](z1, None, None, z1.shape[0])
                i4 = ops.point_query(L[2].table, zc, 4)
                x = ops.voxelize_forward(z2, i4, ops.spcount(i4, L[2].n))
        y2, y4 = ys[1], ys[3]
        if self.spv:
            z3 = ops.devoxelize_forward(y4, *q1) + self.mlps[2](z2, None, None, z2.shape[0])
            cat = torch.cat([z1[:, :self.head_dims[0]], z2[:, :self.head_dims[1]], z3[:, :self.head_dims[2]]], dim=1)
            if cat.shape[1] % 16:
                cat = torch.nn.functional.pad(cat, (0, pad16(cat.shape[1]) - cat.shape[1]))
            logits = self.head(cat.contiguous(), None, None, cat.shape[0], out_dtype=torch.float32)
        else:
            l16 = self.heads[0](x4, None, None, L[4].n, out_dtype=torch.float32, n_dev=nd[4])
            l4 = self.heads[1](y2, None, None, L[2].n, out_dtype=torch.float32, n_dev=nd[2])
            l1 = self.heads[2](y4, None, None, L[0].n, out_dtype=torch.float32, n_dev=nd[0])
            return ops.devoxelize_multi([L[4].table, L[2].table, L[0].table], [16, 4, 1], [l16, l4, l1], zc,
                                        self.num_class, rows=out_rows)
        logits = logits[:, :self.num_class]
        if out_rows is not None:
            logits = ops.gather_rows(logits.contiguous(), out_rows)
        return logits

    @torch.no_grad()
    def forward_batch(self, batch_dict, return_logit=False, return_tta=False):
        """Same batch_dict in / result dict out as the reference model's eval forward."""
        x = batch_dict[self.model.lidar_key]
        out = self(x.C, x.F)
        return self.model.eval_outputs(batch_dict, x, out, return_logit or return_tta)


def tta_vote(point_logits, save_score: bool = False):
    """Test-time-augmentation vote of the reference's evaluation loop (R/train.py:471-475, :497-503): the per-point
    logits of the `votes` augmented copies of one scan (one entry of `point_predict_logits` each) are SUMMED and the
    arg-max is the prediction, returned as int64 class ids (`taseg_b200.io.write_labels` casts them to the uint32 column of the `.label` dump), or
    the float32 summed scores with `save_score`.  Stays on the tensors' device: one D2H copy of N x 4 bytes per scan
    instead of votes x N x classes x 4."""
    total = point_logits[0].clone() if isinstance(point_logits, (list, tuple)) else point_logits.sum(dim=0)
    if isinstance(point_logits, (list, tuple)):
        for extra in point_logits[1:]:
            total += extra
    if save_score:
        return total.float()
    return total.argmax(dim=1).to(torch.int64)
