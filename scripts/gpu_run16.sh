#!/bin/bash
# round 2, GPU run 16 (1 GPU): "after" captures of select_kernel / widen_kernel, predict_tc_kernel<true>, final default line
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 400 $NCU --set full --import-source on -k regex:select_kernel -c 1 -o gpurun_out/r02p_select python scripts/ncu_scenarios.py select 500000 > /dev/null 2>> gpurun_out/r02p.err
timeout 400 $NCU --set full --import-source on -k regex:widen_kernel -c 1 -o gpurun_out/r02p_widen python scripts/ncu_scenarios.py widen 500000 > /dev/null 2>> gpurun_out/r02p.err
timeout 400 $NCU --set full --import-source on -k regex:predict_tc_kernel -c 1 -o gpurun_out/r02p_predict_ragged python bench.py --steps 1 --warmup 3 --mode holdout --calendars 1000 --no-e2e --no-cpu-baseline --no-traffic --no-others > /dev/null 2>> gpurun_out/r02p.err
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02p_launches_holdout_cal1000.csv python bench.py --steps 3 --warmup 3 --mode holdout --calendars 1000 --no-e2e --no-cpu-baseline --no-traffic --no-others > /dev/null 2>> gpurun_out/r02p.err
ls -la gpurun_out/r02p_*.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02p_bench_default.json 2>> gpurun_out/r02p.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02p_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['max_abs_diff_vs_device_path'], d['e2e']['uint16_ingest']['value'])
print('cpu', d['cpu_baseline']['value'])
for k,v in (d.get('other_configs') or {}).items(): print(k, v)
PY
tail -3 gpurun_out/r02p.err
