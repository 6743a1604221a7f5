#!/bin/bash
# round 2, GPU run 18 (1 GPU): final validation of the tree: smoke, full GPU suite, default bench line, reference arm
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02r_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r02r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -4 gpurun_out/r02r_pytest.log
/usr/bin/time -v python bench.py --steps 20 --warmup 3 > gpurun_out/r02r_bench_default.json 2> gpurun_out/r02r.err
grep -E "Elapsed|Maximum resident" gpurun_out/r02r.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02r_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['max_abs_diff_vs_device_path'], d['e2e']['uint16_ingest']['value'])
print('cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
for k,v in (d.get('other_configs') or {}).items(): print(k, v)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02r_bench_reference.json 2>> gpurun_out/r02r.err; cut -c1-200 gpurun_out/r02r_bench_reference.json
timeout 300 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum -k regex:widen_kernel -c 2 --csv --print-units base --log-file gpurun_out/r02r_widen.csv python scripts/ncu_scenarios.py widen 500000 > /dev/null 2>> gpurun_out/r02r.err
grep -v "^==" gpurun_out/r02r_widen.csv | tail -2 | cut -c1-60,200-
