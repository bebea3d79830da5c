#!/bin/bash
# GPU job: parity tests of the v16 conv kernel, then per-layer table and a bench line
set -x
mkdir -p gpurun_out/c16
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/c16/tests.log 2>&1
tail -3 gpurun_out/c16/tests.log
timeout 600 python tools/layer_table.py > gpurun_out/c16/layers.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c16/bench.json 2> gpurun_out/c16/bench.err
tail -c 1500 gpurun_out/c16/bench.json
TSG_FOLD_SHORTCUT=0 TSG_TC_NSPLIT=1 TSG_TC_PDL=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c16/bench_off.json 2> gpurun_out/c16/bench_off.err
tail -c 600 gpurun_out/c16/bench_off.json
