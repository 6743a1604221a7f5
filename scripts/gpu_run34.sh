#!/bin/bash
# round 2, GPU run 34 (1 GPU): design rows of the downdate prefetched into L1 -- now: cache hints (design rows evict_last, records no_allocate)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-traffic --no-others --no-e2e --no-cpu-baseline --nan-frac 0.02"
M="gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"
for lib in default hints; do
  if [ $lib = default ]; then unset MMF_LIB; else export MMF_LIB=$PWD/tests/_build/libmmf_$lib.so; fi
  timeout 300 $B > gpurun_out/r02ae_$lib.json 2>> gpurun_out/r02ae.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02ae_$lib.json').read().strip().splitlines()[-1])
print('$lib', 'nan2 ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
PY
  timeout 300 ncu --clock-control none --metrics $M -k regex:solve_rows_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02ae_$lib.csv $B --steps 1 > /dev/null 2>> gpurun_out/r02ae.err
  grep -v "^==" gpurun_out/r02ae_$lib.csv | python -c "
import csv,sys
print('   ', ' | '.join(r['Metric Name'].split('__')[-1][:28]+'='+r['Metric Value'] for r in csv.DictReader(sys.stdin)))
"
done
tail -2 gpurun_out/r02ae.err
