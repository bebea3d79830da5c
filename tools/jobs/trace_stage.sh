#!/bin/bash
# GPU job: stage-level time stamps of CTA 0 (trace build) for a few layer classes
mkdir -p gpurun_out/ts
export TSG_LIB=$PWD/taseg_b200/libtaseg_b200_trace.so
for c in ${CASES:-4:256:256 3:256:256 3:128:128 0:96:96}; do
  SAMPLES=4 CASES=$c TSG_TC_DEBUG=128 TRACE_ROWS=${ROWS:-36} N=4 timeout 300 python tools/conv_probe.py > gpurun_out/ts/trace_${c//:/_}.txt 2>&1
  cat gpurun_out/ts/trace_${c//:/_}.txt
done
