#!/bin/bash
mkdir -p gpurun_out/n2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2/bench.json 2> gpurun_out/n2/bench.err
tail -c 2500 gpurun_out/n2/bench.json; tail -5 gpurun_out/n2/bench.err
