#!/usr/bin/env python3
"""Throughput of the other BASELINE.json configurations on one GPU (device-resident raw points -> per-point logits),
one JSON line each:
  configs[2]  SPVCNN mk18 cr1.0, single SemanticKITTI-shaped scan per GPU
  configs[3]  TASeg MinkUNetMs mk34 cr1.0 (IN_FEATURE_DIM 4, 17 classes), nuScenes shape: 10 sweeps, 0.1 m voxels, 2 samples per GPU
Scans are independent and the path has no collective, so N GPUs run N copies of this (bench.py measures that scaling on
configs[1]).  Same timing rules as bench.py: warm-up, CUDA events, two batches in flight.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taseg_b200 import frontend, synth  # noqa: E402
from taseg_b200.engine import Engine  # noqa: E402
from taseg_b200.segmentor import SPVCNN, MinkUNetMs, ModelCfg  # noqa: E402


def randomize_bn(model):
    g = torch.Generator().manual_seed(1)
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    return model.cuda().eval()


def timed(step, steps, warmup):
    streams = [torch.cuda.Stream() for _ in range(2)]
    for i in range(warmup):
        with torch.cuda.stream(streams[i % 2]):
            step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        with torch.cuda.stream(streams[i % 2]):
            step()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    args = ap.parse_args()

    # ---- configs[3]: nuScenes shape
    torch.manual_seed(0)
    cfg = ModelCfg(IN_FEATURE_DIM=4, BLOCK="ResBlock", NUM_LAYER=[2, 3, 4, 6, 2, 2, 2, 2], cr=1.0,
                   PLANES=[32, 32, 64, 128, 256, 256, 128, 96, 96], pres=0.1, vres=0.1, IF_DIST=False, IGNORE_LABEL=0, DROPOUT_P=0.0)
    eng = Engine(randomize_bn(MinkUNetMs(cfg, 17)))
    batch = [synth.nus_sample(3000 + b, 10) for b in range(2)]
    dev = [[torch.from_numpy(s).cuda() for s in smp[0]] for smp in batch]
    Rs, Ts, dts = [s[1] for s in batch], [s[2] for s in batch], [s[3] for s in batch]
    info = {}

    def step_nus():
        out = frontend.aggregate_voxelize_nus(dev, Rs, Ts, dts, 0.1)
        info.update(points=int(out["point_ms"].shape[0]), voxels=int(out["coords"].shape[0]), cur=int(sum(out["n_cur"])))
        return eng(out["coords"], out["feats"], out_rows=out["cur_rows"])

    ms = timed(step_nus, args.steps, args.warmup)
    print(json.dumps({"config": "configs[3]: TASeg MinkUNetMs mk34 cr1.0 (IN_FEATURE_DIM 4, 17 classes), nuScenes shape, 10 sweeps, "
                                "0.1 m voxels, batch 2 per GPU", "metric": "scans/sec", "value": 2 / (ms * 1e-3), "ms_per_step": ms,
                      "n_gpus": 1, "dtype": "bf16", "data": "synthetic", **info}))

    # ---- configs[2]: SPVCNN, single frame
    torch.manual_seed(0)
    cfg = ModelCfg(IN_FEATURE_DIM=4, BLOCK="ResBlock", NUM_LAYER=[2, 2, 2, 2, 2, 2, 2, 2], cr=1.0,
                   PLANES=[32, 32, 64, 128, 256, 256, 128, 96, 96], pres=0.05, vres=0.05, IF_DIST=False, IGNORE_LABEL=0, DROPOUT_P=0.0)
    eng2 = Engine(randomize_bn(SPVCNN(cfg, 20)))
    frames, poses = synth.kitti_sample(2000, 1)
    mfb = frontend.MultiFrameBatch([frames], [poses])
    pts = torch.from_numpy(mfb.points).cuda()
    cur_idx = torch.from_numpy(mfb.cur_idx).cuda()
    info2 = {}

    def step_spv():
        out = frontend.aggregate_voxelize(pts, mfb, 0.05, cur_idx)
        info2.update(points=int(out["point_ms"].shape[0]), voxels=int(out["coords"].shape[0]))
        return eng2(out["coords"], out["feats"][:, :4].contiguous())

    ms = timed(step_spv, args.steps, args.warmup)
    print(json.dumps({"config": "configs[2]: SPVCNN mk18 cr1.0, single SemanticKITTI-shaped scan, batch 1 per GPU",
                      "metric": "scans/sec", "value": 1 / (ms * 1e-3), "ms_per_step": ms, "n_gpus": 1, "dtype": "bf16",
                      "data": "synthetic", **info2}))


if __name__ == "__main__":
    main()
