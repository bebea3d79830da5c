#!/usr/bin/env python3
"""Full-size golden for BASELINE configs[1] (what bench.py times): the four 3-frame SemanticKITTI-shaped bench samples
(seeds 2000..2003) through the reference path — numpy multi-scan fusion + clamp/round/shift/sparse_quantize (data
oracle, pinned against the reference's own functions by fuse_kat.npz) and MinkUNetMs mk34 cr1.0 on the reference's
COMPILED torchsparse CPU backend (oracle/_ref via RefOps; the net walker is pinned against the unmodified reference
model classes by net_minkunet_ms.npz).  One scan at a time (the reference CPU kernel_hash mishandles batch > 0).

Run once in the build container (~3 min per sample on 8 cores); writes tests/golden/full_cfg1.npz:
  n_vox[b], n_cur[b], coords_sha[b] (sha256 of the int32 (M,3) voxel list), logits_b (every STEP-th current-scan point).
The GPU test (tests/test_gpu_nets.py::test_full_size_batch4_against_reference) regenerates the same samples from the
seeds and compares the batch-4 engine / fp32 module path with these rows.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
STEP = 97     # prime stride over the current-scan points


def main():
    import bench
    from oracle import data_oracle as D
    from oracle import net_oracle as N
    from oracle import ref_backend as RB
    from oracle import ts_oracle as T
    from taseg_b200 import synth
    assert RB.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    torch.set_num_threads(os.cpu_count() or 1)
    model = bench.make_model("cpu")
    net = N.Net({k: v.numpy() for k, v in model.state_dict().items()}, ops=RB.RefOps)
    out = {"step": np.int64(STEP), "seeds": np.arange(2000, 2000 + bench.BATCH)}
    for b in range(bench.BATCH):
        t0 = time.time()
        frames, poses = synth.kitti_sample(2000 + b, bench.N_FRAMES)
        ms, n0 = D.aggregate_kitti(frames, poses)
        q = D.quantize_ms(ms[:n0], ms, bench.VOXEL)
        coords, feats = T.sparse_collate([q["pc_ms"]], [q["feat_ms"]])
        logits = net.minkunet_ms(coords, feats)
        pts = logits[q["inverse_map_ms"]][:n0]
        out[f"n_vox_{b}"] = np.int64(len(coords))
        out[f"n_cur_{b}"] = np.int64(n0)
        out[f"coords_sha_{b}"] = np.array(hashlib.sha256(np.ascontiguousarray(q["pc_ms"].astype(np.int32)).tobytes()).hexdigest())
        out[f"logits_{b}"] = pts[::STEP].astype(np.float32)
        print("sample %d: %d voxels, %d current points, %.0f s" % (b, len(coords), n0, time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "full_cfg1.npz"), **out)


if __name__ == "__main__":
    main()
