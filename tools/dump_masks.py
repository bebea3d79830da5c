#!/usr/bin/env python3
"""Dump the (tile, offset) activity masks of every pyramid level of one benchmark batch (mask-sorted and lexicographic
tile rows) to gpurun_out/masks.npz — input of tools/sched_sim.py (offline scheduling studies of the tensor-core conv)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from taseg_b200 import frontend  # noqa: E402
from taseg_b200.engine import Geometry  # noqa: E402

samples = bench.make_samples(2000, int(os.environ.get("BATCH", bench.BATCH)))
mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
out = frontend.aggregate_voxelize(torch.from_numpy(mfb.points).cuda(), mfb, bench.VOXEL, torch.from_numpy(mfb.cur_idx).cuda())
geo = Geometry(out["coords"], field_bits=out["field_bits"])
d = {}
for i, lv in enumerate(geo.levels):
    nbr, mask, perm = lv.km3.sorted()
    d["sorted%d" % i] = mask.cpu().numpy()
    d["lex%d" % i] = lv.km3.tile_mask().cpu().numpy()
    d["n%d" % i] = np.int64(lv.n)
    cnt = (nbr[:, :lv.n] >= 0).sum(0) if nbr.shape[1] >= lv.n else None
    print("level", i, "rows", lv.n, "tiles", mask.numel())
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/masks.npz", **d)
