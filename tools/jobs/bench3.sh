#!/bin/bash
mkdir -p gpurun_out/b3
for n in 2 3 4; do
TSG_BENCH_STREAMS=$n timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b3/bench_${n}s.json 2> gpurun_out/b3/bench_${n}s.err
done
