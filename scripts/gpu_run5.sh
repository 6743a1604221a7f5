#!/bin/bash
# round 2, GPU run 5 (1 GPU): full GPU suite (ragged, slabs, integer ingest), multi-destination / ragged bench lines,
# launch lists for small batches, ncu --set full captures of the kernels round 1 had no capture for
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -4 gpurun_out/r02e_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic"
for r in 2 4 8; do timeout 300 $B --replicas $r > gpurun_out/r02e_bench_replicas$r.json 2>> gpurun_out/r02e.err; done
for c in 1 100 1000; do timeout 400 $B --calendars $c > gpurun_out/r02e_bench_cal$c.json 2>> gpurun_out/r02e.err; done
timeout 300 $B --nan-frac 0.02 --calendars 1000 > gpurun_out/r02e_bench_cal1000_nan2.json 2>> gpurun_out/r02e.err
timeout 300 $B --mode holdout > gpurun_out/r02e_bench_holdout.json 2>> gpurun_out/r02e.err
timeout 300 $B --series 10000000 --t 365 > gpurun_out/r02e_bench_cfg5.json 2>> gpurun_out/r02e.err
for f in gpurun_out/r02e_bench_*.json; do python - <<PY
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print('$f'.split('r02e_bench_')[1], 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), d['config'].get('ragged_plan_seconds'))
except Exception as e:
    print('$f', 'FAILED', e)
PY
done
NCU="ncu --clock-control none --profile-from-start off"
# launch lists (durations only) of the timed region at small sizes and for the ragged / gap / holdout steps
for s in 10000 100000; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02e_launches_$s.csv python bench.py --steps 5 --warmup 3 --series $s --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
done
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02e_launches_default.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02e_launches_cal1000.csv python bench.py --steps 3 --warmup 3 --calendars 1000 --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
# full captures
timeout 400 $NCU --set full --import-source on -k regex:fit_tc_kernel -c 1 -o gpurun_out/r02e_fit_tc python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
timeout 400 $NCU --set full --import-source on -k regex:fit_tc_kernel -c 1 -o gpurun_out/r02e_fit_tc_ragged python bench.py --steps 1 --warmup 3 --calendars 1000 --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
timeout 400 $NCU --set full --import-source on -k regex:fit_warp_kernel -c 1 -o gpurun_out/r02e_fit_warp python bench.py --steps 1 --warmup 3 --kernel warp --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
timeout 400 $NCU --set full --import-source on -k "regex:solve_rows_kernel|fit_tc_kernel" -c 2 -o gpurun_out/r02e_nan2 python bench.py --steps 1 --warmup 3 --nan-frac 0.02 --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
timeout 400 $NCU --set full --import-source on -k regex:predict_tc_kernel -c 1 -o gpurun_out/r02e_predict python bench.py --steps 1 --warmup 3 --mode holdout --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02e.err
timeout 400 $NCU --set full --import-source on -k regex:select_kernel -c 1 -o gpurun_out/r02e_select python scripts/ncu_scenarios.py select 500000 > /dev/null 2>> gpurun_out/r02e.err
timeout 400 $NCU --set full --import-source on -k regex:widen_kernel -c 1 -o gpurun_out/r02e_widen python scripts/ncu_scenarios.py widen 500000 > /dev/null 2>> gpurun_out/r02e.err
timeout 500 $NCU --set full --import-source on -k "regex:hash_i32_kernel|verify_i32_kernel|minmax_kernel|scatter_kernel|fill_nan_kernel" -c 5 -o gpurun_out/r02e_packer python scripts/ncu_scenarios.py packer 100000 > /dev/null 2>> gpurun_out/r02e.err
ls -la gpurun_out/*.ncu-rep
tail -5 gpurun_out/r02e.err
