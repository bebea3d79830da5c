#!/usr/bin/env python3
"""Achieved algorithmic bandwidth of the memory-bound kernels of one benchmark step, from an ncu launch list
(`--metrics gpu__time_duration.sum`) and the workload's sizes.  Algorithmic bytes per unit are the ones of DESIGN.md §4.

    python tools/hbm_table.py profiles/launches_r01_final_step.csv > profiles/hbm_kernels_r01.md
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

# bench.py configs[1] workload, batch 4 x 3 frames (profiles/bench_r01_final.json, profiles/conv_layers_r01_v15_step.txt)
N_RAW = 1424247                                      # raw points per step
N_CUR = 474704                                       # current-scan points (outputs)
LEVELS = [702142, 319746, 127983, 46289, 15307]      # voxels at strides 1..16
K3, K2 = 27, 8

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
t = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("tsg::", "").split("<")[0]
    t[name][0] += 1
    t[name][1] += v

sort_keys = (N_RAW * 4 + sum(LEVELS[:4]) * 4            # voxel dedup + 4 down-samples: 4 passes of 9 bits each
             + sum(LEVELS) * 3                           # 3x3x3 mask sorts: 27-bit keys, 3 passes
             + sum(LEVELS[1:]) * 1 + sum(LEVELS[:4]) * 1)  # 2x2x2 down maps (coarse rows) and their transposes (fine rows): 1 pass
rows = [
    ("agg_warp_kernel", "HBM", N_RAW * (16 + 20), "16 B in + 20 B out per point"),
    ("agg_quant_kernel", "HBM", N_RAW * (20 + 16 + 1), "20 B row + 16 B coord + 1 B flag per point"),
    ("agg_shift_kernel", "HBM", N_RAW * 32, "16 B coord read + written per point"),
    ("cp_write_kernel", "HBM", N_RAW * (37 + 36 + 4), "flag + 36 B rows read, 36 B written, 4 B position per point"),
    ("make_coord_keys_kernel", "HBM", (N_RAW + sum(LEVELS[:4])) * 24, "16 B coord read + 8 B key written per row"),
    ("rs_global_hist_kernel", "HBM", (N_RAW + sum(LEVELS[:4]) + sum(LEVELS) + sum(LEVELS[1:]) + sum(LEVELS[:4])) * 8, "8 B key read once per sort"),
    ("rs_pass_kernel", "HBM", sort_keys * 24, "12 B read + 12 B written per key and pass"),
    ("uq_write_kernel", "HBM", (N_RAW + sum(LEVELS[:4])) * (12 + 16 + 8), "12 B read + 16 B coord + first + inverse per key"),
    ("table_insert_coords_kernel", "L2 atomics", sum(LEVELS) * 32, "16 B coord + 16 B slot per voxel"),
    ("table_clear_kernel", "HBM", sum(LEVELS) * 2 * 16, "2 slots x 16 B per voxel"),
    ("kmap_build_kernel", "L2 random 16-B probes", sum(LEVELS) * (16 + K3 * 20) + sum(LEVELS[1:]) * (16 + K2 * 20),
     "16 B coord + K x (16 B probe + 4 B written) per output voxel"),
    ("row_mask_keys_kernel", "HBM", (sum(LEVELS) * K3 + (sum(LEVELS[1:]) + sum(LEVELS[:4])) * K2) * 4 + (2 * sum(LEVELS) + sum(LEVELS[1:4]) * 2) * 8,
     "K x 4 B read + 8 B key per row"),
    ("permute_nbr_kernel", "L2 random 4-B gathers", (sum(LEVELS) * K3 + (sum(LEVELS[1:]) + sum(LEVELS[:4])) * K2) * 8, "K x (4 B gathered + 4 B written) per row"),
    ("devoxelize_multi_kernel", "L2 gathers", N_CUR * (16 + 17 * 16 + 17 * 128 + 80), "coords + 17 probes + 17 rows of 128 B + 80 B written per point"),
    ("cast_pad_bf16_kernel", "HBM", LEVELS[0] * (20 + 32), "20 B read + 32 B written per voxel"),
    ("gather_rows_kernel", "HBM", LEVELS[0] * (4 + 40), "index + 20 B row read + written per voxel"),
]
print("# Memory-bound kernels of one step: algorithmic bytes / ncu time (%s)\n" % os.path.basename(sys.argv[1]))
print("Peak = measured copy bandwidth %.0f GB/s (`MEASURED_PEAKS.json`).  Times are cold-cache, serialised ncu launches, so these" % PEAK)
print("are lower bounds of what the kernels reach inside the pipelined step.  Bytes are ALGORITHMIC (DESIGN.md §4), not measured traffic.\n")
print("| kernel | launches | us | algorithmic MB | GB/s | % of HBM peak | bound | bytes per unit |")
print("|---|---|---|---|---|---|---|---|")
for name, bound, nbytes, note in rows:
    if name not in t:
        continue
    n, us = t[name]
    gbs = nbytes / us / 1e3
    print("| `%s` | %d | %.1f | %.1f | %.0f | %.1f %% | %s | %s |" % (name, n, us, nbytes / 1e6, gbs, 100 * gbs / PEAK, bound, note))
print("\nReading: the single-launch streaming kernels of the front end run at 25-50 % of the copy bandwidth at these sizes (50-100 MB,")
print("10-40 us: launch ramp and tail are a third of the time); the radix-sort passes and the per-level kernels of the small pyramid")
print("levels (46 k and 15 k voxels) are launch-latency bound (10-14 us per launch whatever the size); kernel-map construction, the")
print("row permutation and the devoxelisation tail are bound by random L2 sector traffic (each 4-16 B item costs a 32-B sector), not by HBM.")
