#!/usr/bin/env python3
"""Per-launch table of the tensor-core convolution inside one benchmark step (bench.py's configs[1] workload):
K, channels, rows, pairs, CUDA-event time and algorithmic TFLOP/s of every conv_tc launch, plus totals per layer class.

    python tools/layer_table.py > gpurun_out/layers.txt
"""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from taseg_b200 import frontend, ops  # noqa: E402
from taseg_b200.engine import Engine  # noqa: E402


def main():
    batch = int(os.environ.get("BATCH", bench.BATCH))
    engine = Engine(bench.make_model())
    samples = bench.make_samples(2000, batch)
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    pts = torch.from_numpy(mfb.points).cuda()
    cur_idx = torch.from_numpy(mfb.cur_idx).cuda()

    from taseg_b200.pipeline import Pipeline
    pipe = Pipeline(engine, mfb, bench.VOXEL)
    pipe.calibrate(pts)
    pipe.points.copy_(pts)

    def step():      # sync-free forward, kernel by kernel; the device-side sleep lets the host run ahead of the GPU
        if ops.PROFILE is not None:
            torch.cuda._sleep(int(0.04 * 1.9e9))
        return pipe._forward()

    for _ in range(3):
        step()
    reps = int(os.environ.get("REPS", "5"))
    runs = []
    for _ in range(reps):
        ops.PROFILE = []
        torch.cuda.synchronize()
        step()
        torch.cuda.synchronize()
        runs.append(ops.PROFILE)
        ops.PROFILE = None
    n = len(runs[0])
    rows = []
    for i in range(n):
        k, _, _, pairs, cin, cout, n_out = runs[0][i]
        us = min(r[i][1].elapsed_time(r[i][2]) for r in runs) * 1e3
        rows.append((i, k, cin, cout, int(n_out.item()) if hasattr(n_out, "item") else n_out, int(pairs.item()), us))
    tot_us = sum(r[-1] for r in rows)
    tot_fl = sum(2.0 * r[5] * r[2] * r[3] for r in rows)
    print("# one step = batch %d; %d conv_tc launches; %.1f us total; %.1f GFLOP; %.1f TFLOP/s algorithmic"
          % (batch, n, tot_us, tot_fl / 1e9, tot_fl / tot_us / 1e6))
    print("%3s %3s %4s %4s %8s %9s %9s %8s" % ("i", "K", "cin", "cout", "rows", "pairs", "us", "TFLOP/s"))
    cls = defaultdict(lambda: [0, 0.0, 0.0])
    for i, k, cin, cout, n_out, pairs, us in rows:
        fl = 2.0 * pairs * cin * cout
        print("%3d %3d %4d %4d %8d %9d %9.1f %8.1f" % (i, k, cin, cout, n_out, pairs, us, fl / us / 1e6))
        c = cls[(k, cin, cout, n_out)]
        c[0] += 1
        c[1] += us
        c[2] += fl
    print("\n# by layer class, sorted by time")
    print("%3s %4s %4s %8s %5s %9s %6s %8s" % ("K", "cin", "cout", "rows", "count", "us", "share", "TFLOP/s"))
    for (k, cin, cout, n_out), (cnt, us, fl) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
        print("%3d %4d %4d %8d %5d %9.1f %5.1f%% %8.1f" % (k, cin, cout, n_out, cnt, us, 100 * us / tot_us, fl / us / 1e6))


if __name__ == "__main__":
    main()
