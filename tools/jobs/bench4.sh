#!/bin/bash
mkdir -p gpurun_out/b4
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -5
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/b4/bench.json 2> gpurun_out/b4/bench.err
tail -c 3000 gpurun_out/b4/bench.json; tail -5 gpurun_out/b4/bench.err
