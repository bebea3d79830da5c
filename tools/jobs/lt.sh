#!/bin/bash
# GPU job: per-layer table of the current build (optionally A/B over an environment switch: AB="VAR=a VAR=b")
mkdir -p gpurun_out/lt
i=0
for kv in ${AB:-X=0}; do
  env $kv timeout 300 python tools/layer_table.py > gpurun_out/lt/layers_$i.txt 2>&1
  echo "== $kv"; head -1 gpurun_out/lt/layers_$i.txt; grep -A${ROWS:-12} "by layer class" gpurun_out/lt/layers_$i.txt | tail -n +3
  i=$((i+1))
done
