#!/usr/bin/env python3
"""Timing of the weight-gradient kernels on one benchmark-shaped level (debug aid): LEVEL=0 CIN=96 COUT=96."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from taseg_b200 import frontend, ops  # noqa: E402
from taseg_b200.engine import Geometry  # noqa: E402

samples = bench.make_samples(2000, 4)
mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
out = frontend.aggregate_voxelize(torch.from_numpy(mfb.points).cuda(), mfb, bench.VOXEL, torch.from_numpy(mfb.cur_idx).cuda())
geo = Geometry(out["coords"], field_bits=out["field_bits"])


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


for spec in os.environ.get("CASES", "0:96:96,0:32:32,2:128:128,4:256:256").split(","):
    level, cin, cout = (int(v) for v in spec.split(":"))
    lv = geo.levels[level]
    x = torch.randn(lv.n, cin, device="cuda").bfloat16()
    gy = torch.randn(lv.n, cout, device="cuda").bfloat16()
    pairs = int((lv.km3.nbr >= 0).sum())
    t_list = timed(lambda: ops.PairList(lv.km3.nbr))
    res = []
    for dbg in os.environ.get("DBGS", "0").split(","):
        os.environ["TSG_WG_DEBUG"] = dbg
        res.append("dbg %s: %.0f us" % (dbg, timed(lambda: ops.conv_wgrad_tc(x, gy, lv.km3.nbr, 27))))
    if os.environ.get("PROF"):
        import ctypes
        import numpy as np
        from taseg_b200 import _lib
        os.environ["TSG_WG_DEBUG"] = "64"
        ops.conv_wgrad_tc(x, gy, lv.km3.nbr, 27)
        buf = np.zeros(8, np.int64)
        _lib.lib().tsg_debug_wgrad_prof(buf.ctypes.data_as(ctypes.c_void_p))
        print("   CTA 0: chunks %d | gather warp 0: wait-empty %d, issue %d, loop total %d cycles | MMA: wait-full %d, issue %d"
              % (buf[4], buf[0], buf[1], buf[5], buf[2], buf[3]))
    os.environ["TSG_WG_DEBUG"] = "0"
    t_old = timed(lambda: ops.conv_wgrad_bf16(x, gy, lv.km3.nbr, 27))
    print("level %d %3d->%3d rows %d pairs %d | pair list %.0f us | tcgen05 %s | wmma %.0f us | %.1f TFLOP/s" % (
        level, cin, cout, lv.n, pairs, t_list, ", ".join(res), t_old, 2.0 * pairs * cin * cout / float(res[0].split()[-2]) / 1e6))
