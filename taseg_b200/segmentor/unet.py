"""MinkUNet / MinkUNetMs / SPVCNN backbones (eval + train forward) behind the reference's batch_dict interface.

Graphs:  R/pcseg/model/segmentor/voxel/minkunet/minkunet.py:385-457      (MinkUNet, single frame)
         R/pcseg/model/segmentor/voxel/minkunet/minkunet_ms.py:385-457   (MinkUNetMs, TASeg multi-frame)
         R/pcseg/model/segmentor/fusion/spvcnn/spvcnn.py:399-480         (SPVCNN, point-voxel)
Attribute names (stem, stage1..4, up1..4, classifier, point_transforms) and therefore state_dict keys are the
reference's.  `model_cfgs` is anything with attribute access and .get (the reference passes an EasyDict).
The loss (pcseg.loss.Losses, stock torch, out of scope) is injected as `criterion(logits, target)`.
"""
from typing import Callable, Optional

import torch
from torch import nn

from .. import ops
from ..operators import cat
from ..tensor import PointTensor
from ..utils.lazy import LazyScalar
from .blocks import BLOCKS, BasicConvolutionBlock, BasicDeconvolutionBlock, norm
from .utils import initial_voxelize, point_to_voxel, voxel_to_point
from .. import nn as spnn

__all__ = ['MinkUNet', 'MinkUNetMs', 'MinkUNetMsKd', 'SPVCNN', 'ModelCfg']

DEFAULT_PLANES = [32, 32, 64, 128, 256, 256, 128, 96, 96]


class ModelCfg(dict):
    """Minimal EasyDict stand-in: attribute access + .get, like the reference's model_cfgs."""
    __getattr__ = dict.__getitem__


class _SparseUNet(nn.Module):
    point_branch = False        # SPVCNN adds point MLPs and re-voxelisation
    voxelize_input = True       # MinkUNetMs feeds the de-duplicated voxels directly
    lidar_key, inverse_key = 'lidar', 'inverse_map'

    def __init__(self, model_cfgs, num_class: int, criterion: Optional[Callable] = None):
        super().__init__()
        self.model_cfgs, self.num_class = model_cfgs, num_class
        self.in_feature_dim = model_cfgs.IN_FEATURE_DIM
        self.num_layer = model_cfgs.get('NUM_LAYER', [2, 3, 4, 6, 2, 2, 2, 2])
        self.block = BLOCKS[model_cfgs.get('BLOCK', 'Bottleneck')]
        cr = model_cfgs.get('cr', 1.0)
        cs = [int(cr * c) for c in model_cfgs.get('PLANES', DEFAULT_PLANES)]
        self.pres, self.vres = model_cfgs.get('pres', 0.05), model_cfgs.get('vres', 0.05)
        dist = bool(model_cfgs.IF_DIST)
        exp = self.block.expansion

        self.stem = nn.Sequential(spnn.Conv3d(self.in_feature_dim, cs[0], kernel_size=3, stride=1), norm(cs[0], dist),
                                  spnn.ReLU(True),
                                  spnn.Conv3d(cs[0], cs[0], kernel_size=3, stride=1), norm(cs[0], dist), spnn.ReLU(True))
        width = cs[0]

        def residual_stack(inc, outc, count):
            layers, c = [], inc
            for _ in range(count):
                layers.append(self.block(c, outc, if_dist=dist))
                c = outc * exp
            return layers, c

        for i in range(4):                                  # encoder: stride-2 conv then residual blocks
            blocks, out_w = residual_stack(width, cs[1 + i], self.num_layer[i])
            setattr(self, f'stage{i + 1}', nn.Sequential(
                BasicConvolutionBlock(width, width, ks=2, stride=2, dilation=1, if_dist=dist), *blocks))
            width = out_w
        skips = [cs[3] * exp, cs[2] * exp, cs[1] * exp, cs[0]]
        for i in range(4):                                  # decoder: transposed conv, concat skip, residual blocks
            blocks, out_w = residual_stack(cs[5 + i] + skips[i], cs[5 + i], self.num_layer[4 + i])
            setattr(self, f'up{i + 1}', nn.ModuleList([
                BasicDeconvolutionBlock(width, cs[5 + i], ks=2, stride=2, if_dist=dist), nn.Sequential(*blocks)]))
            width = out_w
        self.classifier = nn.Sequential(nn.Linear((cs[4] + cs[6] + cs[8]) * exp, num_class))
        if self.point_branch:
            dims = [cs[0], cs[4] * exp, cs[6] * exp, cs[8] * exp]
            self.point_transforms = nn.ModuleList([
                nn.Sequential(nn.Linear(dims[i], dims[i + 1]),
                              nn.SyncBatchNorm(dims[i + 1]) if dist else nn.BatchNorm1d(dims[i + 1]), nn.ReLU(True))
                for i in range(3)])
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.SyncBatchNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.dropout = nn.Dropout(model_cfgs.get('DROPOUT_P', 0.3), True)
        self.criterion = criterion

    # reference checkpoint loader (R/pcseg/model/segmentor/base_segmentors.py:16-27)
    def load_params(self, model_state_disk, strict=False):
        mine = self.state_dict()
        picked = {}
        for key, value in model_state_disk.items():
            key = key[len('module.'):] if key.startswith('module.') else key
            if key in mine and mine[key].shape == value.shape:
                picked[key] = value
        return self.load_state_dict(picked, strict=strict)

    def features(self, x, z, sfx: str = ''):
        """x: input SparseTensor (voxels), z: PointTensor -> the concatenated multi-scale point features the classifier
        reads.  `sfx` selects a second set of backbone modules (`stem_gt`, ... of the distillation model)."""
        m = (lambda name: getattr(self, name + sfx)) if sfx else (lambda name: getattr(self, name))
        spv = self.point_branch
        x0 = m('stem')(x)
        z0 = voxel_to_point(x0, z, nearest=False)
        x1 = m('stage1')(point_to_voxel(x0, z0) if spv else x0)
        x2 = m('stage2')(x1)
        x3 = m('stage3')(x2)
        x4 = m('stage4')(x3)
        z1 = voxel_to_point(x4, z0)
        if spv:
            z1.F = z1.F + self.point_transforms[0](z0.F)
            y1 = point_to_voxel(x4, z1)
        else:
            y1 = x4
        y1.F = m('dropout')(y1.F)
        y1 = m('up1')[1](cat([m('up1')[0](y1), x3]))
        y2 = m('up2')[1](cat([m('up2')[0](y1), x2]))
        z2 = voxel_to_point(y2, z1)
        if spv:
            z2.F = z2.F + self.point_transforms[1](z1.F)
            y3 = point_to_voxel(y2, z2)
        else:
            y3 = y2
        y3.F = m('dropout')(y3.F)
        y3 = m('up3')[1](cat([m('up3')[0](y3), x1]))
        y4 = m('up4')[1](cat([m('up4')[0](y3), x0]))
        z3 = voxel_to_point(y4, z2)
        if spv:
            z3.F = z3.F + self.point_transforms[2](z2.F)
        return torch.cat([z1.F, z2.F, z3.F], dim=1)

    def backbone(self, x, z):
        """x: input SparseTensor (voxels), z: PointTensor -> per-point (or per-voxel for Ms) logits."""
        return self.classifier(self.features(x, z))

    def logits(self, lidar):
        lidar.F = lidar.F[:, :self.in_feature_dim]
        z = PointTensor(lidar.F, lidar.C.float())
        x = initial_voxelize(z, self.pres, self.vres) if self.voxelize_input else lidar
        return self.backbone(x, z)

    def forward(self, batch_dict, return_logit=False, return_tta=False):
        x = batch_dict[self.lidar_key]
        out = self.logits(x)
        if self.training:
            key = 'targets_ms' if self.lidar_key == 'lidar_ms' else 'targets'
            target = batch_dict[key].F.long().cuda(non_blocking=True)
            crit = self.criterion or (lambda o, t: nn.functional.cross_entropy(o, t, ignore_index=self.model_cfgs.IGNORE_LABEL))
            loss = crit(out, target)
            lazy = LazyScalar(loss)      # read like a float; the host waits for it only when the logger looks (utils/lazy.py)
            return {'loss': loss}, {'loss': lazy}, {'loss': lazy}
        return self.eval_outputs(batch_dict, x, out, return_logit or return_tta)

    def eval_outputs(self, batch_dict, x, out, want_prob):
        """Per-sample scatter of voxel logits back to the raw points of the current scan
        (minkunet.py:434-457 / minkunet_ms.py:434-457), done with device gathers; one D2H copy per sample."""
        invs, labels = batch_dict[self.inverse_key], batch_dict['targets_mapped']
        ms = self.lidar_key == 'lidar_ms'
        vox_b, inv_b, lab_b = x.C[:, -1], invs.C[:, -1], labels.C[:, -1]
        n_batch = int(inv_b.max().item()) + 1
        vox_start = torch.searchsorted(vox_b.contiguous(), torch.arange(n_batch + 1, device=vox_b.device, dtype=vox_b.dtype))
        num_points = [int(v) for v in torch.as_tensor(batch_dict['num_points']).reshape(-1).tolist()]
        pointer = 0
        ret = {'point_predict': [], 'point_labels': [], 'point_predict_logits': [], 'name': batch_dict['name']}
        for b in range(n_batch):
            rows = invs.F[inv_b == b].long() + vox_start[b]
            if ms:
                n_ms = int(torch.as_tensor(batch_dict['num_points_ms']).reshape(-1)[b])
                rows = rows[batch_dict['point_mask'][pointer:pointer + n_ms].to(rows.device)]
                pointer += n_ms
            rows = rows[:num_points[b]]
            logits = ops.gather_rows(out.float(), rows)
            ret['point_predict'].append((logits.softmax(1) if want_prob else logits.argmax(1)).cpu().numpy())
            ret['point_predict_logits'].append(logits.cpu().numpy())
            ret['point_labels'].append(labels.F[lab_b == b][:num_points[b]].cpu().numpy())
        return ret


class MinkUNet(_SparseUNet):
    pass


class MinkUNetMs(_SparseUNet):
    voxelize_input = False
    lidar_key, inverse_key = 'lidar_ms', 'inverse_map_ms'


class MinkUNetMsKd(MinkUNetMs):
    """TASeg's distillation segmentor (R/pcseg/model/segmentor/voxel/minkunet/minkunet_ms_kd.py:200-666): a frozen
    teacher backbone — modules `stem_gt`, `stage1_gt` .. `up4_gt`, `classifier_gt`, `dropout_gt`, fed `lidar_ms_gt` (the
    multi-frame cloud aggregated with ground-truth FSA masks) — next to the student backbone on `lidar_ms`.  Training adds
    an MSE feature-distillation term on the voxels both inputs contain: student rows are matched to teacher rows with
    sphash / sphashquery (:613-615), per sample at most MAX_VOXEL matched voxels are drawn (SAMPLING_TYPE 'random',
    :617-633; other sampling types add nothing in the reference either).  Eval is the student's (:637-666).
    Module names, hence state_dict keys, are the reference's."""
    TWIN = ('stem', 'stage1', 'stage2', 'stage3', 'stage4', 'up1', 'up2', 'up3', 'up4', 'classifier', 'dropout')

    def __init__(self, model_cfgs, num_class: int, criterion: Optional[Callable] = None):
        super().__init__(model_cfgs, num_class, criterion)
        teacher = MinkUNetMs(model_cfgs, num_class)
        for name in self.TWIN:
            setattr(self, name + '_gt', getattr(teacher, name))
        self.sampling_type = model_cfgs.get('SAMPLING_TYPE', 'uncertain')
        self.max_voxel = model_cfgs.get('MAX_VOXEL', 3000)
        self.feat_kd_weight = model_cfgs.get('FEAT_KD_WEIGHT', 1.0)

    def teacher_features(self, batch_dict):
        with torch.no_grad():
            x_gt = batch_dict['lidar_ms_gt']
            x_gt.F = x_gt.F[:, :self.in_feature_dim]
            return x_gt, self.features(x_gt, PointTensor(x_gt.F, x_gt.C.float()), '_gt')

    def distillation_loss(self, x, feat, x_gt, feat_gt):
        """MSE between the student's and the (detached) teacher's point features on the voxels they share."""
        from ..nn import functional as F
        s2d = F.sphashquery(F.sphash(x.C.int()), F.sphash(x_gt.C.int()))
        loss = feat.new_zeros(())
        if self.sampling_type != 'random':
            return loss, s2d
        batch_size = int(x.C[:, -1].max().item()) + 1
        for b in range(batch_size):
            mask = (s2d >= 0) & (x.C[:, -1] == b)
            if int(mask.sum().item()) > self.max_voxel:
                inds = mask.nonzero().reshape(-1)
                mask = mask.clone()
                mask[inds[torch.randperm(len(inds), device=inds.device)][self.max_voxel:]] = False
            loss = loss + nn.functional.mse_loss(feat[mask], feat_gt[s2d[mask]].detach()) * self.feat_kd_weight / batch_size
        return loss, s2d

    def forward(self, batch_dict, return_logit=False, return_tta=False):
        x_gt, feat_gt = self.teacher_features(batch_dict)
        x = batch_dict['lidar_ms']
        x.F = x.F[:, :self.in_feature_dim]
        feat = self.features(x, PointTensor(x.F, x.C.float()))
        out = self.classifier(feat)
        if not self.training:
            return self.eval_outputs(batch_dict, x, out, return_logit or return_tta)
        target = batch_dict['targets_ms'].F.long().cuda(non_blocking=True)
        crit = self.criterion or (lambda o, t: nn.functional.cross_entropy(o, t, ignore_index=self.model_cfgs.IGNORE_LABEL))
        loss_seg = crit(out, target)
        loss_kd, _ = self.distillation_loss(x, feat, x_gt, feat_gt)
        loss = loss_seg + loss_kd
        info = {'loss': loss.item(), 'loss_seg': loss_seg.item(), 'loss_feat_kd': float(loss_kd.detach())}
        return {'loss': loss}, dict(info), dict(info)


class SPVCNN(_SparseUNet):
    point_branch = True
