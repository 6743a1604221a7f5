#!/bin/bash
# round 2, GPU run 30 (1 GPU): occupancy of solve_rows_kernel -- 214 registers x 8 warps/SM against 168 registers x 12 warps/SM
mkdir -p gpurun_out
B="python bench.py --nan-frac 0.02 --steps 20 --warmup 3 --no-traffic --no-others --no-e2e --no-cpu-baseline"
for lib in default s128x3 s96x3 s32x9 default s128x3; do
  if [ $lib = default ]; then unset MMF_LIB; else export MMF_LIB=$PWD/tests/_build/libmmf_$lib.so; fi
  timeout 300 $B > gpurun_out/r02ab_$lib.json 2>> gpurun_out/r02ab.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02ab_$lib.json').read().strip().splitlines()[-1])
print('$lib', 'nan2 ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
PY
done
tail -2 gpurun_out/r02ab.err
