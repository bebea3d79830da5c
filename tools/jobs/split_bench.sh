#!/bin/bash
# GPU job: bench.py value / e2e / conv_ms with the K split off and on (same box)
mkdir -p gpurun_out/sb
i=0
for cfg in "TSG_SPLIT_K=0" "TSG_SPLIT_K=1 TSG_SPLIT_CAP=14 TSG_SPLIT_PARTS=2" "TSG_SPLIT_K=1 TSG_SPLIT_CAP=10 TSG_SPLIT_PARTS=3" "TSG_SPLIT_K=0" "TSG_SPLIT_K=1 TSG_SPLIT_CAP=14 TSG_SPLIT_PARTS=2"; do
  env $cfg timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/sb/b$i.json 2> gpurun_out/sb/b$i.err
  python - "$cfg" $i <<'PY'
import json, sys
d = json.loads(open('gpurun_out/sb/b%s.json' % sys.argv[2]).read().strip().split('\n')[-1])
print('%-50s value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.4f' % (sys.argv[1], d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac']))
PY
  i=$((i+1))
done
