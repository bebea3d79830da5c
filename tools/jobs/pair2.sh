#!/bin/bash
# GPU job: CTA pairs with two relay warps — parity of the op tests in pair mode, then the per-layer table with pairs off / on
mkdir -p gpurun_out/p2
TSG_TC_PAIR=1 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_tensor_core" 2>&1 | tail -3
for m in 0 1 0 1; do
  TSG_TC_PAIR=$m timeout 300 python tools/layer_table.py > gpurun_out/p2/layers_pair${m}.txt 2>&1
  echo "== TSG_TC_PAIR=$m"; head -1 gpurun_out/p2/layers_pair$m.txt
  grep -E "^ +27 +(256|128|384|192|64) +(256|128) +[0-9]+ +[0-9]+ " gpurun_out/p2/layers_pair$m.txt
done
