#!/bin/bash
# round 2, GPU run 43 (2-GPU box, ONE process): the default bench line as the driver's N=1 step of a scaling run would see it
mkdir -p gpurun_out
timeout 600 python bench.py --no-traffic --no-others > gpurun_out/r02ai_bench_1proc_2gpubox.json 2> gpurun_out/r02ai.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ai_bench_1proc_2gpubox.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['host_narrow'], '|', e['transport'][:60], '| u16', e.get('uint16_ingest',{}).get('value'))
print('cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
PY
tail -2 gpurun_out/r02ai.err
