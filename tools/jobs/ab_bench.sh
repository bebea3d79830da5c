#!/bin/bash
# GPU job: headline bench under A/B switches: AB="A=1,B=2 A=0,B=2" (comma-separated env per variant)
mkdir -p gpurun_out/ab
i=0
for v in $AB; do
  env $(echo $v | tr ',' ' ') timeout 600 python bench.py --steps ${STEPS:-30} --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ab/b$i.json 2> gpurun_out/ab/b$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab/b$i.json"))
print("$v", "value %.1f e2e %.1f ms %.3f conv_ms %.3f frac %.3f share %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"],d["roofline"]["kernel_share_of_step"]))
PY
  i=$((i+1))
done
