#!/bin/bash
# GPU job: ncu --set full of every layer class of the shipped conv kernel + clock64 hand-off traces (trace build)
set -x
mkdir -p gpurun_out/d1
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "point_voxel or conv_tensor_core" > gpurun_out/d1/tests.log 2>&1
SAMPLES=4 ONLY=sorted timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/d1/prof_v15 python tools/profile_conv.py > gpurun_out/d1/ncu.log 2>&1
export TSG_LIB=$PWD/taseg_b200/libtaseg_b200_trace.so
C=0:96:96,0:32:32,3:128:128,4:256:256,2:64:64,3:256:256
SAMPLES=4 CASES=$C TSG_TC_DEBUG=128 N=6 TRACE_ROWS=48 timeout 400 python tools/conv_probe.py > gpurun_out/d1/trace_stage.txt 2>&1
SAMPLES=4 CASES=$C TSG_TC_DEBUG=128 TRACE_TILES=1 N=6 timeout 400 python tools/conv_probe.py > gpurun_out/d1/trace_tile.txt 2>&1
SAMPLES=4 CASES=3:128:128,0:32:32,0:96:96,4:256:256 DBGS=0,1,2,4,8,3,15 N=8 timeout 400 python tools/conv_probe.py > gpurun_out/d1/knockout.txt 2>&1
ls -la gpurun_out/d1
