#!/usr/bin/env python3
"""Host-side (Python) cost of one benchmark step: cProfile over a few steps, top functions by own time."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from taseg_b200 import frontend  # noqa: E402
from taseg_b200.engine import Engine  # noqa: E402

engine = Engine(bench.make_model())
samples = bench.make_samples(2000, bench.BATCH)
mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
pts = torch.from_numpy(mfb.points).cuda()
cur_idx = torch.from_numpy(mfb.cur_idx).cuda()


def step():
    out = frontend.aggregate_voxelize(pts, mfb, bench.VOXEL, cur_idx)
    return engine(out["coords"], out["feats"], field_bits=out["field_bits"], out_rows=out["cur_rows"])


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue time per step %.3f ms, with final sync %.3f ms" % ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
