#!/bin/bash
# GPU job: v18 conv (CTA pairs): parity on the op tests, then the per-layer table with pairs off / default / everywhere
mkdir -p gpurun_out/v18
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_tensor_core" > gpurun_out/v18/tests.log 2>&1
tail -15 gpurun_out/v18/tests.log
for m in 1 2 0; do
  TSG_TC_PAIR=$m timeout 300 python tools/layer_table.py > gpurun_out/v18/layers_pair$m.txt 2>&1
  echo "== TSG_TC_PAIR=$m"; head -1 gpurun_out/v18/layers_pair$m.txt; tail -3 gpurun_out/v18/layers_pair$m.txt
done
