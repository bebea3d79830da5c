#!/bin/bash
# GPU job: batches in flight (value / e2e), same box
mkdir -p gpurun_out/ss
for n in ${NS:-4 5 6 7}; do
  TSG_BENCH_STREAMS=$n timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ss/s$n.json 2> gpurun_out/ss/s$n.err
  python - $n <<'PY'
import json, sys
d = json.loads(open('gpurun_out/ss/s%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
print('streams %s value %.1f (%.3f ms) e2e %.1f (%.3f ms)' % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
PY
done
