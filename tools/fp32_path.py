#!/usr/bin/env python3
"""Speed of the fp32 drop-in (module) path at benchmark size: forward of MinkUNetMs mk34 cr1.0 on one batch of 4 three-frame scans,
eval mode, fp32 tensors — what `train.py` without --amp and fp32 evaluation through `pcseg` get.  TSG_FP32_SPLIT=0/1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import taseg_b200 as ts  # noqa: E402
from taseg_b200 import frontend  # noqa: E402

model = bench.make_model()
samples = bench.make_samples(2000, 4)
mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
out = frontend.aggregate_voxelize(torch.from_numpy(mfb.points).cuda(), mfb, bench.VOXEL, torch.from_numpy(mfb.cur_idx).cuda())
coords, feats = out["coords"], out["feats"]


def fwd():
    with torch.no_grad():
        x = ts.SparseTensor(feats.clone(), coords)
        return model.logits(x)


try:
    fwd()
except Exception as e:  # noqa
    print("forward failed:", repr(e)[:300])
    raise
torch.cuda.synchronize()
ts_ = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fwd()
    e1.record()
    torch.cuda.synchronize()
    ts_.append(e0.elapsed_time(e1))
print("fp32 module-path forward, batch 4 (%d voxels), TSG_FP32_SPLIT=%s: %.1f ms" % (coords.shape[0], os.environ.get("TSG_FP32_SPLIT", "1"), min(ts_)))
