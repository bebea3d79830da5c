#!/usr/bin/env python3
"""Time the geometry kernels (tables, kernel maps, mask sorts, uniques) of one benchmark-shaped batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from taseg_b200 import frontend, ops  # noqa: E402
from taseg_b200.nn.utils.kernel import kernel_offsets_np  # noqa: E402


_burn = torch.randn(4096, 4096, device="cuda")


def timed(fn, reps=5):
    fn()
    for _ in range(20):      # keep the clocks up: the GPU idles while the host prepares each case
        _burn @ _burn
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, r


samples = bench.make_samples(2000, int(os.environ.get("BATCH", "4")))
mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
pts = torch.from_numpy(mfb.points).cuda()
cur_idx = torch.from_numpy(mfb.cur_idx).cuda()
us, out = timed(lambda: frontend.aggregate_voxelize(pts, mfb, bench.VOXEL, cur_idx))
print("front end (aggregate + quantise + dedup): %.1f us" % us)
c = out["coords"]
for l in range(5):
    s = 2 ** l
    us_t, table = timed(lambda: ops.Table.from_coords(c))
    us_k, km3 = timed(lambda: ops.build_kmap(table, c.shape[0], c, kernel_offsets_np(3, s)))
    us_s, _ = timed(lambda: ops.KernelMap(km3.nbr, km3.nbsizes32, km3.blockcnt, km3.n_in, km3.n_out, 27).sorted())
    line = "level %d n=%7d: table %.1f us, kmap3 %.1f us, sort rows %.1f us" % (l, c.shape[0], us_t, us_k, us_s)
    if l < 4:
        us_u, nxt = timed(lambda: ops.unique_coords(c, trunc_stride=2 * s, field_bits=out["field_bits"]))
        us_k2, km2 = timed(lambda: ops.build_kmap(table, c.shape[0], nxt, kernel_offsets_np(2, s)))
        line += ", downsample %.1f us, kmap2 %.1f us" % (us_u, us_k2)
        c = nxt
    print(line)
