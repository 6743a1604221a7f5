#!/bin/bash
# round 2, GPU run 40 (1 GPU): final-tree ncu evidence -- launch list of the default bench command, --set full of fit_tc_kernel
# (gap-free) and of fit_tc_kernel + solve_rows_kernel (2 % missing)
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02ag_launches_default.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic --no-others > /dev/null 2>> gpurun_out/r02ag.err
grep -v "^==" gpurun_out/r02ag_launches_default.csv | tail -12 | cut -c1-200
timeout 400 $NCU --set full --import-source on -k regex:fit_tc_kernel -c 1 -o gpurun_out/r02ag_fit_tc python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic --no-others > /dev/null 2>> gpurun_out/r02ag.err
timeout 400 $NCU --set full --import-source on -k "regex:solve_rows_kernel|fit_tc_kernel" -c 2 -o gpurun_out/r02ag_nan2 python bench.py --steps 1 --warmup 3 --nan-frac 0.02 --no-e2e --no-cpu-baseline --no-traffic --no-others > /dev/null 2>> gpurun_out/r02ag.err
ls -la gpurun_out/r02ag_*.ncu-rep; tail -2 gpurun_out/r02ag.err
