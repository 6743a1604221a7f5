#!/bin/bash
# round 2, GPU run 15 (1 GPU): ragged holdout (predict_tc MULTI) + full GPU suite + memcheck of the new path
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -25 gpurun_out/r02o_pytest.log
timeout 900 compute-sanitizer --error-exitcode 9 --tool memcheck python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k "ragged" > gpurun_out/r02o_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02o_memcheck.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic --no-others"
for c in 1 1000; do timeout 400 $B --mode holdout --calendars $c > gpurun_out/r02o_bench_holdout_cal$c.json 2>> gpurun_out/r02o.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02o_bench_holdout_cal$c.json').read().strip().splitlines()[-1])
    print('holdout cal$c', 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), d['config'].get('ragged_plan_seconds'))
except Exception as e: print('holdout cal$c FAILED', e)
PY
done
tail -3 gpurun_out/r02o.err
