#!/bin/bash
# round 2, GPU run 28 (1 GPU): default bench line with the added "2 % of the series have a 10-day gap" entry
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02aa_bench_default.json 2> gpurun_out/r02aa.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02aa_bench_default.json').read().strip().splitlines()[-1])
print('default ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d['other_configs'].items(): print(' ', k, v if isinstance(v,str) else (round(v['ms_per_step'],4), round(v['roofline_frac'],3)))
PY
tail -2 gpurun_out/r02aa.err
