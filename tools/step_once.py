#!/usr/bin/env python3
"""One benchmark step (bench.py's configs[1] workload) bracketed by cudaProfilerStart/Stop, for ncu launch lists:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/step_once.py

BATCH / FRAMES env vars shrink the workload for `--set full` captures.
"""
import os
import sys

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

    if os.environ.get("TSG_BENCH_EAGER") == "1":
        def step():
            out = frontend.aggregate_voxelize(pts, mfb, bench.VOXEL, cur_idx)
            return engine(out["coords"], out["feats"], field_bits=out["field_bits"], out_rows=out["cur_rows"])
    else:      # the shipped path: sync-free pipeline, captured graph (GRAPH=0: same kernels launched one by one)
        from taseg_b200.pipeline import Pipeline
        pipe = Pipeline(engine, mfb, bench.VOXEL)
        pipe.calibrate(pts)
        pipe.points.copy_(pts)
        if os.environ.get("GRAPH", "1") == "1":
            pipe.capture()

        def step():
            return pipe()

    for _ in range(int(os.environ.get("WARM", "2"))):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
