#!/usr/bin/env python3
"""Per-CTA timeline of one tensor-core convolution launch (trace build, TSG_TC_DEBUG=128): when every CTA started and
finished (globaltimer), how many stages / tiles it processed.  Answers: ramp, tail, imbalance, cycles per stage."""
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
for spec in os.environ.get("CASES", "3:128:128,3:256:256,4:256:256,2:64:64,2:128:128,0:96:96").split(","):
    level, cin, cout = (int(v) for v in spec.split(":"))
    lv = geo.levels[level]
    x = torch.randn(lv.n, cin, device="cuda").bfloat16()
    packed = ops.pack_weights(torch.randn(27, cin, cout, device="cuda") * 0.05, cin)
    nbr, mask, perm = lv.km3.sorted()
    split = lv.km3.split_items() if os.environ.get("SPLIT") else None
    for _ in range(3):
        ops.conv_forward_tc(x, None, packed, 27, cout, nbr, mask, lv.n, perm=perm, relu=True, split=split)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_forward_tc(x, None, packed, 27, cout, nbr, mask, lv.n, perm=perm, relu=True, split=split)
    e1.record()
    torch.cuda.synchronize()
    buf = np.zeros((160, 4), np.int64)
    _lib.lib().tsg_debug_conv_ctas(buf.ctypes.data_as(ctypes.c_void_p))
    b = buf[buf[:, 1] > 0]
    t0 = b[:, 0].min()
    start, end = (b[:, 0] - t0) / 1e3, (b[:, 1] - t0) / 1e3
    dur = end - start
    print("level %d %3d->%3d rows %d: event %.1f us | %d CTAs, start spread %.1f us, end min/median/max %.1f/%.1f/%.1f us, busy median %.1f us | "
          "stages(MMA warp 0) min/median/max %d/%d/%d, tiles %d/%d/%d | ns per stage (busy/2*stages) median %.0f"
          % (level, cin, cout, lv.n, e0.elapsed_time(e1) * 1e3, len(b), start.max(), end.min(), np.median(end), end.max(), np.median(dur),
             b[:, 2].min(), np.median(b[:, 2]), b[:, 2].max(), b[:, 3].min(), np.median(b[:, 3]), b[:, 3].max(),
             np.median(dur * 1e3 / np.maximum(2 * b[:, 2], 1))))
    order = np.argsort(end)
    print("   slowest CTAs (id: start end stages tiles):", " ".join("%d:%.1f-%.1f/%d/%d" % (i, start[i], end[i], b[i, 2], b[i, 3]) for i in order[-6:]))
    if os.environ.get("ALL"):
        print("   all CTAs by id:", " ".join("%d:%.0f-%.0f/%d/%d" % (i, start[i], end[i], b[i, 2], b[i, 3]) for i in range(len(b))))
    print("   fastest CTAs:", " ".join("%d:%.1f-%.1f/%d/%d" % (i, start[i], end[i], b[i, 2], b[i, 3]) for i in order[:6]))
