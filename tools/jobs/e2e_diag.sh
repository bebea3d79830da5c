#!/bin/bash
# GPU job: which part of the end-to-end step costs the ~5 % between `value` and `e2e` (diagnostic switches, no e2e claim printed)
mkdir -p gpurun_out/e2e
for sk in ${SKIPS:-none d2h h2d,d2h,stage,out}; do
  TSG_BENCH_E2E_SKIP=$([ $sk = none ] && echo "" || echo $sk) timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/e2e/$sk.json 2> gpurun_out/e2e/$sk.err
  python - "$sk" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/e2e/%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
print('skip %-22s value ms/step %.3f   e2e ms/step %.3f' % (sys.argv[1], d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done
