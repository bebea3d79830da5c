#!/bin/bash
mkdir -p gpurun_out/v18
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_tensor_core" > gpurun_out/v18/tests_b.log 2>&1
tail -3 gpurun_out/v18/tests_b.log
TSG_LIB=$PWD/taseg_b200/libtaseg_b200_trace.so TSG_TC_DEBUG=128 timeout 300 python tools/cta_timeline.py 2>&1 | grep "^level"
timeout 300 python tools/layer_table.py > gpurun_out/v18/layers_throttle.txt 2>&1
head -1 gpurun_out/v18/layers_throttle.txt; grep -A32 "by layer class" gpurun_out/v18/layers_throttle.txt
