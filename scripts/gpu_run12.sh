#!/bin/bash
# round 2, GPU run 12 (1 GPU): final-build validation: smoke, full GPU suite, default bench line (with other_configs),
# compute-sanitizer over the new code paths
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02l_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -4 gpurun_out/r02l_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02l_bench_default.json 2> gpurun_out/r02l.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02l_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['max_abs_diff_vs_device_path'], d['e2e']['uint16_ingest']['value'])
print('cpu', d['cpu_baseline']['value'])
print('others', json.dumps(d.get('other_configs'), indent=1))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02l_bench_reference.json 2>> gpurun_out/r02l.err; cut -c1-300 gpurun_out/r02l_bench_reference.json
CS="compute-sanitizer --error-exitcode 9"
timeout 900 $CS --tool memcheck python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k "ragged_calendars or integer_ingest or streaming or narrowing or captured or select" > gpurun_out/r02l_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02l_memcheck.log
timeout 900 $CS --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/r02l_racecheck.log
timeout 600 $CS --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r02l_synccheck.log
tail -3 gpurun_out/r02l.err
