#!/bin/bash
# GPU job: n-way K split — test, then the per-layer table of the step with the split off / on at a few caps
mkdir -p gpurun_out/sp
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "k_split or tensor_core" 2>&1 | tail -3
i=0
for cfg in "TSG_SPLIT_K=0" "TSG_SPLIT_K=1 TSG_SPLIT_CAP=7 TSG_SPLIT_PARTS=4" "TSG_SPLIT_K=1 TSG_SPLIT_CAP=9 TSG_SPLIT_PARTS=3" "TSG_SPLIT_K=1 TSG_SPLIT_CAP=5 TSG_SPLIT_PARTS=6" "TSG_SPLIT_K=1 TSG_SPLIT_CAP=14 TSG_SPLIT_PARTS=2"; do
  env $cfg timeout 300 python tools/layer_table.py > gpurun_out/sp/layers_$i.txt 2>&1
  echo "== $cfg"; head -1 gpurun_out/sp/layers_$i.txt; grep -E "^ +27 +256 +256 +15307|^ +27 +128 +256 +15307|^ +8 +128 +128 +15307" gpurun_out/sp/layers_$i.txt | tail -4
  i=$((i+1))
done
