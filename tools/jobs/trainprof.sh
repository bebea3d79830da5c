#!/bin/bash
mkdir -p gpurun_out/tp
PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tp/train_launches.csv python tools/train_bench.py --batch 4 --steps 1 --warmup 2 > gpurun_out/tp/train.log 2>&1
tail -3 gpurun_out/tp/train.log
python tools/launch_summary.py gpurun_out/tp/train_launches.csv | head -50
