#!/bin/bash
# round 2, GPU run 32 (1 GPU): solve_rows_kernel sorts each warp's 128 work-list entries by gap count, one flat downdate loop
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02ad_pytest.log 2>&1; tail -4 gpurun_out/r02ad_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-traffic --no-others --no-e2e --no-cpu-baseline"
for nf in 0.02 0.002 0.1; do
  timeout 300 $B --nan-frac $nf > gpurun_out/r02ad_nan_$nf.json 2>> gpurun_out/r02ad.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02ad_nan_$nf.json').read().strip().splitlines()[-1])
print('nan-frac', $nf, 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
PY
done
timeout 300 $B --nan-frac 0.02 --mode holdout > gpurun_out/r02ad_nan_holdout.json 2>> gpurun_out/r02ad.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ad_nan_holdout.json').read().strip().splitlines()[-1])
print('holdout nan2 ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
PY
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed -k regex:solve_rows_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02ad_solve.csv $B --nan-frac 0.02 --steps 1 > /dev/null 2>> gpurun_out/r02ad.err
grep -v "^==" gpurun_out/r02ad_solve.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print('  ', r['Metric Name'], r['Metric Value'])
"
tail -2 gpurun_out/r02ad.err
