#!/bin/bash
# GPU job: is the gather pipeline latency-bound or throughput-bound?  Sweep the number of stages / G on the trace build.
mkdir -p gpurun_out/st
export TSG_LIB=$PWD/taseg_b200/libtaseg_b200_trace.so
C=0:96:96,0:32:32,3:128:128,4:256:256,2:64:64,3:256:256,2:128:128
for st in 8 3 2; do
  echo "== stages<=$st" 
  SAMPLES=4 CASES=$C TSG_TC_STAGES=$st N=8 timeout 300 python tools/conv_probe.py 2>&1 | grep level
done
echo "== G1"
for st in 8 4 3; do
  echo "== G1 stages<=$st"
  SAMPLES=4 CASES=0:96:96,0:32:32,2:64:64,2:128:128 TSG_TC_G1=1 TSG_TC_STAGES=$st N=8 timeout 300 python tools/conv_probe.py 2>&1 | grep level
done
