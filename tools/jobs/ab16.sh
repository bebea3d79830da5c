#!/bin/bash
set -x
mkdir -p gpurun_out/ab16
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/ab16/$name.json 2> gpurun_out/ab16/$name.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/ab16/$name.json').read().strip().split('\n')[-1])
print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_step'],3))
"; }
run fold_only TSG_TC_NSPLIT=1 TSG_TC_PDL=0
run fold_pdl TSG_TC_NSPLIT=1 TSG_TC_PDL=1
run fold_ns TSG_TC_PDL=0
run all_off TSG_FOLD_SHORTCUT=0 TSG_TC_NSPLIT=1 TSG_TC_PDL=0
run fold_only_1stream TSG_TC_NSPLIT=1 TSG_TC_PDL=0 TSG_BENCH_STREAMS=1
run fold_pdl_1stream TSG_TC_NSPLIT=1 TSG_TC_PDL=1 TSG_BENCH_STREAMS=1
