#!/usr/bin/env python3
"""BASELINE.json configs[4]: TASeg MinkUNetMs training step (bf16 autocast, SGD) with the NCCL weight-gradient all-reduce.

    python tools/train_bench.py [--steps 5] [--batch 2]                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...

One process per GPU, every rank trains on its own synthetic 3-frame samples (weak scaling); the only collective is the
bucketed gradient all-reduce of taseg_b200.parallel.GradientReducer, overlapped with backward.  Prints one JSON line
(rank 0): ms per step (CUDA events, max over ranks), scans/s, loss, gradient bytes per step, and a parameter checksum
spread over ranks (0 = replicas identical after the steps).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import taseg_b200 as ts  # noqa: E402
from taseg_b200 import frontend, parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--sync-bn", action="store_true", help="IF_DIST=True: SyncBatchNorm (batch statistics over all ranks, R/pcseg/model/segmentor/voxel/minkunet/minkunet.py:23-25)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = bench.make_model(if_dist=args.sync_bn and world > 1).train()
    samples = bench.make_samples(5000 + rank * args.batch, args.batch)
    mfb = frontend.MultiFrameBatch([s[0] for s in samples], [s[1] for s in samples])
    pts = torch.from_numpy(mfb.points).cuda()
    cur_idx = torch.from_numpy(mfb.cur_idx).cuda()
    out = frontend.aggregate_voxelize(pts, mfb, bench.VOXEL, cur_idx)
    coords, feats = out["coords"], out["feats"]
    g = torch.Generator(device="cuda").manual_seed(rank)
    labels = torch.randint(1, 20, (coords.shape[0],), device="cuda", generator=g)
    opt = torch.optim.SGD(model.parameters(), lr=0.02, momentum=0.9, weight_decay=1e-4)
    reducer = parallel.GradientReducer(model.parameters(), bucket_mb=25.0)

    def step():
        batch = {"lidar_ms": ts.SparseTensor(feats.clone(), coords), "targets_ms": ts.SparseTensor(labels, coords)}
        return parallel.train_step(model, batch, opt, reducer, amp_dtype=torch.bfloat16, sync=os.environ.get("TRAIN_SYNC", "0") == "1")

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if os.environ.get("PROFILE_STEP"):     # ncu --profile-from-start off: the launch list of exactly one training step
        torch.cuda.cudart().cudaProfilerStart()
        loss = step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    if os.environ.get("HOST_PROFILE"):     # where the host spends its time enqueuing a step (cProfile, host time only)
        import cProfile
        import pstats
        import time
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        for _ in range(5):
            loss = step()
        pr.disable()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("host enqueue %.2f ms per step, drain %.2f ms after the last enqueue" % ((t1 - t0) * 200, (t2 - t1) * 1e3))
        pstats.Stats(pr).sort_stats("tottime").print_stats(int(os.environ["HOST_PROFILE"]))
        pstats.Stats(pr).sort_stats("cumulative").print_stats(int(os.environ["HOST_PROFILE"]))
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / args.steps, "cuda")
    check = torch.stack([torch.stack([p.detach().double().sum() for p in model.parameters()]).sum(),
                         torch.stack([b.double().sum() for b in model.buffers() if b.dtype.is_floating_point]).sum()])
    spread, bn_spread = 0.0, 0.0      # parameters / BN running statistics (the latter differ by design without --sync-bn)
    if world > 1:
        allc = [torch.zeros_like(check) for _ in range(world)]
        dist.all_gather(allc, check)
        allc = torch.stack(allc)
        spread, bn_spread = [float(v) for v in (allc.max(dim=0).values - allc.min(dim=0).values).abs()]
    if rank == 0:
        print(json.dumps({"metric": "train step (MinkUNetMs mk34 cr1.0, 3-frame KITTI shape, bf16 autocast)", "n_gpus": world,
                          "batch_per_gpu": args.batch, "sync_bn": bool(args.sync_bn and world > 1), "ms_per_step": ms, "scans_per_s": args.batch * world / (ms * 1e-3),
                          "loss": float(loss), "voxels_per_gpu": int(coords.shape[0]),
                          "allreduce_bytes_per_step": reducer.bytes_per_step(), "buckets": len(reducer.buckets),
                          "param_checksum_spread_over_ranks": spread, "bn_running_stat_spread_over_ranks": bn_spread,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
