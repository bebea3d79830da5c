#!/usr/bin/env python3
"""What a tensor-core convolution launch costs before and after its main loop (trace build, TSG_TC_DEBUG=128):
CUDA-event time of tiny launches (1x1 and 3x3x3 on the stride-16 level) and CTA 0's clock64 life-cycle stamps."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from taseg_b200 import _lib, frontend, ops  # noqa: E402
from taseg_b200.engine import Geometry  # noqa: E402

samples = bench.make_samples(2000, 4)
mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
out = frontend.aggregate_voxelize(torch.from_numpy(mfb.points).cuda(), mfb, bench.VOXEL, torch.from_numpy(mfb.cur_idx).cuda())
geo = Geometry(out["coords"], field_bits=out["field_bits"])


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def life():
    buf = np.zeros(8, np.int64)
    _lib.lib().tsg_debug_conv_life(buf.ctypes.data_as(ctypes.c_void_p))
    return (buf[1:6] - buf[0]).tolist()


for level, cin, cout, k in [(4, 256, 32, 1), (4, 256, 256, 1), (4, 128, 128, 27), (4, 256, 256, 27), (3, 128, 128, 27), (2, 64, 64, 27)]:
    lv = geo.levels[level]
    x = torch.randn(lv.n, cin, device="cuda").bfloat16()
    packed = ops.pack_weights(torch.randn(k, cin, cout, device="cuda") * 0.05, cin)
    if k == 1:
        args = dict(nbr=None, tile_mask=None, perm=None)
    else:
        nbr, mask, perm = lv.km3.sorted()
        args = dict(nbr=nbr, tile_mask=mask, perm=perm)
    for split in ([None] if k == 1 else [None, lv.km3.split_items()]):
        if k != 1 and split is None and False:
            continue
        us = timed(lambda: ops.conv_forward_tc(x, None, packed, k, cout, args["nbr"], args["tile_mask"], lv.n, perm=args["perm"], relu=True, split=split))
        print("level %d rows %6d K %2d %3d->%3d split %d: %6.1f us  life (cycles after entry: prologue done, 1st plan, producers see plan, "
              "last epilogue, exit) %s" % (level, lv.n, k, cin, cout, split is not None, us, life() if os.environ.get("TSG_TC_DEBUG") else ""))
