#!/usr/bin/env python3
"""tests/golden/kd.npz: the UNMODIFIED reference MinkUNetMsKd (R/pcseg/model/segmentor/voxel/minkunet/minkunet_ms_kd.py)
in training mode on the compiled CPU backend — student / teacher point features, the sphashquery matching and the
feature-distillation loss (SAMPLING_TYPE 'random', no sub-sampling), for tests/test_gpu_nets.py::test_kd_twin_backbone.
Run in the build container (needs /root/reference and oracle/_ref):  python tests/golden/make_golden_kd.py"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    G.import_reference()
    from oracle import data_oracle as D
    from torchsparse import SparseTensor
    from torchsparse.utils.collate import sparse_collate
    from pcseg.model.segmentor.voxel.minkunet.minkunet_ms_kd import MinkUNetMsKd
    frames, poses = G.small_sample(7100, 3)
    ms, n0 = D.aggregate_kitti(frames, poses)
    rng = np.random.default_rng(5)
    keep = np.ones(len(ms), bool)
    keep[n0:] = rng.random(len(ms) - n0) < 0.5            # the student sees half of the history points
    qt = D.quantize_ms(ms[:n0], ms, 0.05)
    qs = D.quantize_ms(ms[:n0], ms[keep], 0.05)
    cfg = G.model_cfg(5, 0.125, [1, 1, 1, 1, 1, 1, 1, 1])
    cfg.update(SAMPLING_TYPE="random", MAX_VOXEL=10 ** 9, FEAT_KD="mse", FEAT_KD_WEIGHT=2.0)
    torch.manual_seed(0)
    model = MinkUNetMsKd(cfg, 20)
    G.randomize_bn(model, 1)
    model.train()
    sd = {k: v.clone() for k, v in model.state_dict().items()}      # before the step updates the BN running statistics
    lidar_t = sparse_collate([SparseTensor(torch.from_numpy(qt["feat_ms"]), torch.from_numpy(qt["pc_ms"]))])
    lidar_s = sparse_collate([SparseTensor(torch.from_numpy(qs["feat_ms"]), torch.from_numpy(qs["pc_ms"]))])
    labels = torch.from_numpy(rng.integers(1, 20, len(qs["pc_ms"])).astype(np.int64))
    cap = {}

    class FeatKd(torch.nn.Module):          # records what the reference hands to its MSELoss, returns the same losses
        def forward(self, ys, yt):
            cap["s"], cap["t"] = ys[0].detach().clone(), yt[0].detach().clone()
            return [torch.nn.functional.mse_loss(a, b) for a, b in zip(ys, yt)]

    class NoSegLoss(torch.nn.Module):       # pcseg.loss.Losses (CE + Lovasz) is outside the hot path
        def forward(self, out, target, **kw):
            return out.sum() * 0.0
    model.criterion_feat_kd = FeatKd()
    model.criterion_losses = NoSegLoss()
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self                 # the forward calls .cuda() on the targets
    try:
        ret, tb, disp = model(dict(lidar_ms_gt=lidar_t, lidar_ms=lidar_s, targets_ms=SparseTensor(labels, lidar_s.C),
                                   offset_ms=torch.tensor([len(labels)])))
    finally:
        torch.Tensor.cuda = orig_cuda
    import torchsparse.nn.functional as F
    s2d = F.sphashquery(F.sphash(lidar_s.C.int()), F.sphash(lidar_t.C.int())).numpy()
    matched = s2d >= 0
    assert cap["s"].shape[0] == int(matched.sum())
    out = dict(coords_s=lidar_s.C.numpy(), feats_s=lidar_s.F.numpy(), coords_t=lidar_t.C.numpy(), feats_t=lidar_t.F.numpy(),
               s2d_sha=np.array(hashlib.sha256(np.ascontiguousarray(s2d.astype(np.int64)).tobytes()).hexdigest()),
               n_matched=np.array(int(matched.sum())), feat_s=cap["s"].numpy()[::8], feat_t=cap["t"].numpy()[::8],
               loss_feat_kd=np.array(tb["loss_feat_kd"], np.float64))
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "kd.npz"), **out)
    print("kd.npz: student voxels %d, teacher voxels %d, matched %d, loss_feat_kd %.6f, %d state entries" %
          (len(lidar_s.C), len(lidar_t.C), int(matched.sum()), tb["loss_feat_kd"], len(sd)))


if __name__ == "__main__":
    main()
