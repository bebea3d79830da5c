#!/bin/bash
mkdir -p gpurun_out/pv
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/pv/tests.log 2>&1
tail -25 gpurun_out/pv/tests.log
timeout 600 python tools/pv_bench.py > gpurun_out/pv/pv_bench.md 2> gpurun_out/pv/pv_bench.err
cat gpurun_out/pv/pv_bench.md; tail -5 gpurun_out/pv/pv_bench.err
