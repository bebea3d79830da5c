#!/usr/bin/env python3
"""Per-launch timing distribution of one tensor-core convolution shape (debug aid).
    LEVEL=0 CIN=96 COUT=96 N=20 python tools/conv_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taseg_b200 import frontend, ops, synth  # noqa: E402
from taseg_b200.engine import Geometry  # noqa: E402

samples = [synth.kitti_sample(2000 + i, 3) for i in range(int(os.environ.get("SAMPLES", "2")))]
mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
out = frontend.aggregate_voxelize(torch.from_numpy(mfb.points).cuda(), mfb, 0.05, torch.from_numpy(mfb.cur_idx).cuda())
geo = Geometry(out["coords"], field_bits=out["field_bits"])
for spec in os.environ.get("CASES", "0:96:96").split(","):
    level, cin, cout = (int(v) for v in spec.split(":"))
    lv = geo.levels[level]
    x = torch.randn(lv.n, cin, device="cuda").bfloat16()
    packed = ops.pack_weights(torch.randn(27, cin, cout, device="cuda") * 0.05, cin)
    nbr, mask, perm = lv.km3.sorted()
    res = torch.randn(lv.n, cout, device="cuda").bfloat16() if os.environ.get("RES") else None
    if os.environ.get("NOPERM"):
        perm = None
    for dbg in os.environ.get("DBGS", os.environ.get("TSG_TC_DEBUG", "0")).split(","):   # knock-out sweep (trace build)
        os.environ["TSG_TC_DEBUG"] = dbg
        ts = []
        for i in range(int(os.environ.get("N", "20"))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_forward_tc(x, None, packed, 27, cout, nbr, mask, lv.n, perm=perm, residual=res, relu=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print("level %d %3d->%3d rows %d dbg %s res %d perm %d:" % (level, cin, cout, lv.n, dbg, res is not None, perm is not None),
              " ".join("%.0f" % t for t in ts[-6:]), "us")
    if int(os.environ.get("TSG_TC_DEBUG", "0")) & 128:
        import ctypes
        import numpy as np
        from taseg_b200 import _lib
        buf = np.zeros((13, 96), np.int64)
        _lib.lib().tsg_debug_conv_trace(buf.ctypes.data_as(ctypes.c_void_p))
        if os.environ.get("TRACE_TILES"):
            tb = buf[8:12]
            t0 = tb[tb > 0].min()
            print(" tile stages  M.start    M.end  E.start    E.end | mainloop  epilogue  cycles/stage")
            for i in range(96):
                if buf[8, i] == 0:
                    break
                ms, me, es, ee, ns = (int(buf[r, i]) for r in (8, 9, 10, 11, 12))
                print("%5d %6d %8d %8d %8d %8d | %8d %8d %8.0f" % (i, ns, ms - t0, me - t0, es - t0, ee - t0, me - ms, ee - es,
                                                                  (me - ms) / max(ns, 1)))
            continue
        t0 = buf[buf > 0].min()
        names = ["P0.empty", "P0.arrive", "M.full", "M.commit", "W.empty", "M.wait", "P0.copied", "P0.idxnext"]
        print("stage " + " ".join("%9s" % n for n in names))
        for i in range(int(os.environ.get("TRACE_ROWS", "40"))):
            print("%5d " % i + " ".join("%9d" % (buf[r, i] - t0 if buf[r, i] else -1) for r in range(8)))
