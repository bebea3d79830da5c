#!/usr/bin/env python3
"""Achieved HBM bandwidth of the point <-> voxel transforms at SPVCNN sizes (BASELINE configs[2]: 8 single scans of
~118 k points, 0.05 m voxels; point features of 32..96 channels), against the measured copy bandwidth.

    python tools/pv_bench.py > gpurun_out/pv_bench.md

Algorithmic bytes: voxelize = N rows read + M rows written + 4 N (order) + 8 M (start, count); devoxelize = N rows written +
64 N (8 indices + 8 weights) + gathered voxel rows counted ONCE (M rows: the 8 N gathers hit L2).  The timed region flushes
nothing: inputs of the larger cases exceed the 126 MB L2."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taseg_b200 import ops, synth  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    peak = 6546.2
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    batch = int(os.environ.get("BATCH", "8"))
    pts = []
    for b in range(batch):
        f, _ = synth.kitti_sample(3000 + b, 1)
        p = np.concatenate([f[0][:, :3] / 0.05, np.full((len(f[0]), 1), b, np.float32)], 1).astype(np.float32)
        pts.append(p)
    pc = torch.from_numpy(np.concatenate(pts)).cuda()
    n = pc.shape[0]
    floor_c = torch.cat([torch.floor(pc[:, :3]).int(), pc[:, 3:].int()], 1)
    vox, _, inv = ops.unique_coords(floor_c, want_index=True, want_inverse=True, by_hash=True)
    m = vox.shape[0]
    cnt = ops.spcount(inv, m)
    tab = ops.Table.from_coords(vox.contiguous())
    idx8, w8 = ops.trilinear_query(tab, pc, 1)
    print("# point <-> voxel transforms, %d scans: N = %d points, M = %d voxels; peak = measured copy bandwidth %.0f GB/s\n" % (batch, n, m, peak))
    print("| kernel | dtype | C | us | algorithmic MB | GB/s | % of HBM peak |")
    print("|---|---|---|---|---|---|---|")
    for dtype, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        es = 4 if dtype == torch.float32 else 2
        for c in (32, 64, 96):
            f = torch.randn(n, c, device="cuda").to(dtype)
            vf = torch.randn(m, c, device="cuda").to(dtype)
            plan = ops.VoxelizePlan.of(inv, m)
            t = timeit(lambda: ops.voxelize_forward(f, inv, cnt, plan=plan))
            mb = (n * c * es + m * c * es + 4 * n + 8 * m) / 1e6
            print("| `voxelize_seg_kernel` (scatter-mean) | %s | %d | %.1f | %.1f | %.0f | %.1f %% |" % (name, c, t, mb, mb / t * 1e3, mb / t * 1e5 / peak))
            t = timeit(lambda: ops.devoxelize_forward(vf, idx8, w8))
            mb = (n * c * es + 64 * n + m * c * es) / 1e6
            print("| `devoxelize_vec_kernel` (trilinear) | %s | %d | %.1f | %.1f | %.0f | %.1f %% |" % (name, c, t, mb, mb / t * 1e3, mb / t * 1e5 / peak))
    f = torch.randn(n, 32, device="cuda")
    t = timeit(lambda: ops.VoxelizePlan(inv, m))
    print("\nplan (radix sort of %d point ids by voxel + segment heads): %.1f us, built once per index tensor" % (n, t))
    # the reference-style kernels on the same inputs (scalar, atomics) for comparison
    from taseg_b200 import _lib
    from taseg_b200._lib import call, ptr, stream
    out = torch.empty((m, 32), device="cuda")
    t = timeit(lambda: call("tsg_voxelize_fwd", ptr(f), 0, ptr(inv), ptr(cnt), n, 32, m, ptr(out), None, stream()))
    print("reference-style scalar atomic voxelize (fp32, C = 32): %.1f us" % t)
    vf = torch.randn(m, 32, device="cuda")
    o2 = torch.empty((n, 32), device="cuda")
    t = timeit(lambda: call("tsg_devoxelize_fwd", ptr(vf), 0, ptr(idx8), ptr(w8), n, 32, ptr(o2), stream()))
    print("reference-style scalar devoxelize (fp32, C = 32): %.1f us" % t)


if __name__ == "__main__":
    main()
