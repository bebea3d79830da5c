#!/usr/bin/env python3
"""Run a few representative tensor-core convolution launches on one benchmark-shaped sample (for ncu captures).

    ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 6 -o gpurun_out/prof python tools/profile_conv.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taseg_b200 import frontend, ops, synth  # noqa: E402
from taseg_b200.engine import Geometry  # noqa: E402


def main():
    n_samples = int(os.environ.get("SAMPLES", "2"))
    samples = [synth.kitti_sample(2000 + i, 3) for i in range(n_samples)]
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    out = frontend.aggregate_voxelize(torch.from_numpy(mfb.points).cuda(), mfb, 0.05, torch.from_numpy(mfb.cur_idx).cuda())
    geo = Geometry(out["coords"], field_bits=out["field_bits"])
    torch.manual_seed(0)
    cases = [(0, 96, 96), (1, 32, 32), (0, 32, 32), (2, 64, 64), (4, 256, 256), (3, 128, 128), (3, 256, 256), (2, 128, 128), (1, 96, 96)]
    reps = int(os.environ.get("REPS", "1"))
    if os.environ.get("CASES"):
        cases = [cases[int(i)] for i in os.environ["CASES"].split(",")]
    for level, cin, cout in cases:
        lv = geo.levels[level]
        x = torch.randn(lv.n, cin, device="cuda").bfloat16()
        w = torch.randn(27, cin, cout, device="cuda") * 0.05
        packed = ops.pack_weights(w, cin)
        pairs = int((lv.km3.nbr >= 0).sum())
        tiles = (lv.n + 127) // 128
        for name, (nbr, mask, perm) in (("lex   ", (lv.km3.nbr, lv.km3.tile_mask(), None)), ("sorted", lv.km3.sorted())):
            if os.environ.get("ONLY") and os.environ["ONLY"] != name.strip():
                continue
            active = int(sum(bin(v & 0xffffffff).count("1") for v in mask.tolist()))
            ops.conv_forward_tc(x, None, packed, 27, cout, nbr, mask, lv.n, perm=perm)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                y = ops.conv_forward_tc(x, None, packed, 27, cout, nbr, mask, lv.n, perm=perm)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            dense = 2.0 * active * 128 * cin * cout
            print("level %d (stride %2d) n=%7d pairs=%8d %s active (tile,offset) %.2f  %3d->%3d : %8.1f us  %6.1f TFLOP/s "
                  "algorithmic, %6.1f TFLOP/s issued" % (level, lv.stride, lv.n, pairs, name, active / (27.0 * tiles), cin,
                                                         cout, ms * 1e3, 2.0 * pairs * cin * cout / ms / 1e9, dense / ms / 1e9))


if __name__ == "__main__":
    main()
