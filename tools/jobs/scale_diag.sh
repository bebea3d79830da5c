#!/bin/bash
# GPU job (gpurun --gpus N): the device-resident and end-to-end step at N ranks with per-rank times and host enqueue time
N=${N:-4}
mkdir -p gpurun_out/sd
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/sd/n$N.json 2> gpurun_out/sd/n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open('gpurun_out/sd/n%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
for k in ('value', 'ms_per_step', 'e2e', 'host_enqueue_ms_per_step', 'ms_per_step_by_rank', 'clocks'):
    print(k, d.get(k))
PY
tail -3 gpurun_out/sd/n$N.err
