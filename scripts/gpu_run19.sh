#!/bin/bash
# round 2, GPU run 19 (1 GPU): default bench line on the final tree, timed
mkdir -p gpurun_out
SECONDS=0
python bench.py --steps 20 --warmup 3 > gpurun_out/r02s_bench_default.json 2> gpurun_out/r02s.err
echo "bench wall seconds: $SECONDS"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['max_abs_diff_vs_device_path'], d['e2e']['uint16_ingest']['value'])
print('cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
for k,v in (d.get('other_configs') or {}).items(): print(k, v)
PY
tail -3 gpurun_out/r02s.err
