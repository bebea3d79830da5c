#!/bin/bash
# GPU job: v17 conv (staging swizzle, static first ticket, K-split work items): parity, fixed-cost probe, layer table, bench
mkdir -p gpurun_out/v17
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_tensor_core or kernel_maps" > gpurun_out/v17/tests.log 2>&1
tail -5 gpurun_out/v17/tests.log
timeout 300 python tools/fixed_cost.py > gpurun_out/v17/fixed_prod.txt 2>&1
TSG_LIB=$PWD/taseg_b200/libtaseg_b200_trace.so TSG_TC_DEBUG=128 timeout 300 python tools/fixed_cost.py > gpurun_out/v17/fixed_trace.txt 2>&1
cat gpurun_out/v17/fixed_prod.txt gpurun_out/v17/fixed_trace.txt
timeout 600 python tools/layer_table.py > gpurun_out/v17/layers.txt 2>&1
tail -32 gpurun_out/v17/layers.txt
TSG_SPLIT_K=0 timeout 600 python tools/layer_table.py > gpurun_out/v17/layers_nosplit.txt 2>&1
head -3 gpurun_out/v17/layers_nosplit.txt
