#!/bin/bash
# round 2, GPU run 10 (1 GPU): half-chunk TMEM stores at 80 registers + streaming solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k "streaming or capture or integer" > gpurun_out/r02j_pytest.log 2>&1
tail -3 gpurun_out/r02j_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/r02j_$name.json 2>> gpurun_out/r02j.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02j_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'])
except Exception as e: print('$name FAILED', e)
PY
}
R80=$PWD/tests/_build/libmmf_r80h.so; HALF=$PWD/tests/_build/libmmf_half.so
run default_prod $B
run default_half env MMF_LIB=$HALF $B
run default_r80h env MMF_LIB=$R80 $B
run default_r80h_stream env MMF_LIB=$R80 $B --stream-solve
run nan2_prod $B --nan-frac 0.02
run nan2_prod_stream $B --nan-frac 0.02 --stream-solve
run nan2_r80h env MMF_LIB=$R80 $B --nan-frac 0.02
run nan2_r80h_stream env MMF_LIB=$R80 $B --nan-frac 0.02 --stream-solve
run nan02_r80h_stream env MMF_LIB=$R80 $B --nan-frac 0.002 --stream-solve
run nan02_prod $B --nan-frac 0.002
tail -3 gpurun_out/r02j.err
