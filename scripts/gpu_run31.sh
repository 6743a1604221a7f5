#!/bin/bash
# round 2, GPU run 31 (1 GPU): solve_rows_kernel under ncu at 214 regs (8 warps/SM) and 168 regs (12 warps/SM)
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_registers,sm__maximum_warps_per_active_cycle_pct,launch__grid_size,launch__block_size,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio"
for lib in default s128x3 s32x9; do
  if [ $lib = default ]; then unset MMF_LIB; else export MMF_LIB=$PWD/tests/_build/libmmf_$lib.so; fi
  timeout 400 ncu --clock-control none --metrics $M -k regex:solve_rows_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02ac_$lib.csv python bench.py --steps 1 --warmup 3 --nan-frac 0.02 --no-e2e --no-cpu-baseline --no-traffic --no-others > /dev/null 2>> gpurun_out/r02ac.err
  echo "== $lib"; grep -v "^==" gpurun_out/r02ac_$lib.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print('  ', r['Metric Name'], r['Metric Value'])
"
done
tail -2 gpurun_out/r02ac.err
