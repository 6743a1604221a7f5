#!/bin/bash
# round 2, GPU run 25 (1 GPU): A/B of the balanced sub-wave launch on ONE box, back-to-back steps (other_configs)
mkdir -p gpurun_out
for v in 1 0 1 0; do
  timeout 400 python bench.py --steps 20 --warmup 3 --tc-variant $v --no-traffic > gpurun_out/r02x_v$v.json 2>> gpurun_out/r02x.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02x_v$v.json').read().strip().splitlines()[-1])
o=d['other_configs']
k='configs[1] 10k x 1095'
print('variant', $v, 'C2 ms', round(o[k]['ms_per_step'],5), 'finite', round(o[k]['assume_finite']['ms_per_step'],5), 'C3', round(o['configs[2] 100k x 1095']['ms_per_step'],5), 'default', round(d['ms_per_step'],4))
PY
done
