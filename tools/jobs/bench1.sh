#!/bin/bash
mkdir -p gpurun_out/b1
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b1/bench.json 2> gpurun_out/b1/bench.err
tail -c 2500 gpurun_out/b1/bench.json; tail -5 gpurun_out/b1/bench.err
TSG_BENCH_STREAMS=1 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b1/bench_1s.json 2> gpurun_out/b1/bench_1s.err
tail -c 600 gpurun_out/b1/bench_1s.json
TSG_BENCH_EAGER=1 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b1/bench_eager.json 2> gpurun_out/b1/bench_eager.err
tail -c 600 gpurun_out/b1/bench_eager.json
timeout 600 python tools/layer_table.py > gpurun_out/b1/layers.txt 2>&1
tail -45 gpurun_out/b1/layers.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b1/launches.csv python tools/step_once.py > gpurun_out/b1/ncu.log 2>&1
tail -3 gpurun_out/b1/ncu.log
