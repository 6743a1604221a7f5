#!/bin/bash
# round 2, GPU run 9 (1 GPU): streaming solve beside the tcgen05 kernel; 80-register build vs ptxas' 128
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -5 gpurun_out/r02i_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/r02i_$name.json 2>> gpurun_out/r02i.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02i_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'])
except Exception as e: print('$name FAILED', e)
PY
}
run default_r80 $B
MMF_LIB=$PWD/tests/_build/libmmf_r128.so run default_r128 env MMF_LIB=$PWD/tests/_build/libmmf_r128.so $B
run nan2_r80 $B --nan-frac 0.02
run nan2_r128 env MMF_LIB=$PWD/tests/_build/libmmf_r128.so $B --nan-frac 0.02
run nan02_r80 $B --nan-frac 0.002
run c3_r80 $B --series 100000 --steps 50
run c3_r128 env MMF_LIB=$PWD/tests/_build/libmmf_r128.so $B --series 100000 --steps 50
run c2_r80 $B --series 10000 --steps 50
run holdout_r80 $B --mode holdout
run nan2_holdout_r80 $B --mode holdout --nan-frac 0.02
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02i_launches_nan2.csv python bench.py --steps 3 --warmup 3 --nan-frac 0.02 --no-e2e --no-cpu-baseline --no-traffic > /dev/null 2>> gpurun_out/r02i.err
grep -v "^==" gpurun_out/r02i_launches_nan2.csv | tail -8
tail -3 gpurun_out/r02i.err
