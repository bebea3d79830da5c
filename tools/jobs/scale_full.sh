#!/bin/bash
# GPU job (gpurun --gpus N): the default bench line at N ranks (configs[1] + other_configs)
N=${N:-8}
mkdir -p gpurun_out/sf
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/sf/bench_n$N.json 2> gpurun_out/sf/bench_n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open('gpurun_out/sf/bench_n%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
for k in ('value', 'ms_per_step', 'e2e', 'host_enqueue_ms_per_step', 'ms_per_step_by_rank', 'clocks'):
    print(k, d.get(k))
for k, v in (d.get('other_configs') or {}).items():
    print(k, {a: b for a, b in v.items() if a != 'workload'})
PY
tail -3 gpurun_out/sf/bench_n$N.err
