#!/bin/bash
mkdir -p gpurun_out/fin
timeout 300 python tools/layer_table.py > gpurun_out/fin/layers.txt 2>&1
head -1 gpurun_out/fin/layers.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/fin/bench_n1.json 2> gpurun_out/fin/bench_n1.err
tail -c 1200 gpurun_out/fin/bench_n1.json; tail -2 gpurun_out/fin/bench_n1.err
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')
"
