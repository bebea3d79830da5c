#!/bin/bash
mkdir -p gpurun_out/b2
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b2/bench.json 2> gpurun_out/b2/bench.err
tail -5 gpurun_out/b2/bench.err
TSG_BENCH_STREAMS=3 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b2/bench_3s.json 2> gpurun_out/b2/bench_3s.err
TSG_BENCH_STREAMS=1 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b2/bench_1s.json 2> gpurun_out/b2/bench_1s.err
