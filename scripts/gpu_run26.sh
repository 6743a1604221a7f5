#!/bin/bash
# round 2, GPU run 26 (1 GPU): final tree -- whole GPU suite, smoke, default bench line (live traffic, other configs), reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02y_pytest.log 2>&1; tail -3 gpurun_out/r02y_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02y_bench_default.json').read().strip().splitlines()[-1])
print('default ms', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
for k,v in d['other_configs'].items(): print(' ', k, round(v['ms_per_step'],4), round(v['roofline_frac'],3))
PY
tail -2 gpurun_out/r02y.err
