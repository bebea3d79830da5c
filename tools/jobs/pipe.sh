#!/bin/bash
mkdir -p gpurun_out/p
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/p/tests.log 2>&1
tail -40 gpurun_out/p/tests.log
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -m gpu > gpurun_out/p/tests2.log 2>&1
tail -5 gpurun_out/p/tests2.log
