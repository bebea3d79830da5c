#!/bin/bash
mkdir -p gpurun_out/wg
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "wgrad or vs_oracle" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
for m in 1 0; do
  TSG_WGRAD_TC=$m timeout 600 python tools/train_bench.py --batch 4 --steps 5 --warmup 2 2>&1 | tail -1 | tee gpurun_out/wg/train_tc$m.json
done
PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/wg/train_launches.csv python tools/train_bench.py --batch 4 --steps 1 --warmup 2 > gpurun_out/wg/train.log 2>&1
python tools/launch_summary.py gpurun_out/wg/train_launches.csv 2>/dev/null | head -24
