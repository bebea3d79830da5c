#!/bin/bash
# GPU job: pipeline parity after geometry changes, launch list of one step, A/B bench
mkdir -p gpurun_out/geo
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_nets.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/geo/launches_step.csv python tools/step_once.py > gpurun_out/geo/step.log 2>&1
python tools/launch_summary.py gpurun_out/geo/launches_step.csv 2>/dev/null | head -16
AB="X=0 X=1" STEPS=30 bash tools/jobs/ab_bench.sh
