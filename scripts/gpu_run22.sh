#!/bin/bash
# round 2, GPU run 22 (1 GPU): balanced launches of small batches -- bit-equality test, whole GPU suite, C2/C3/literal-C4 lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k balanced > gpurun_out/r02v_bal.log 2>&1; tail -5 gpurun_out/r02v_bal.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02v_pytest.log 2>&1; tail -3 gpurun_out/r02v_pytest.log
for n in 10000 100000 125000 250000; do
  for v in 1 0; do
    timeout 300 python bench.py --series $n --steps 50 --warmup 5 --tc-variant $v --no-traffic --no-others > gpurun_out/r02v_bench_${n}_v${v}.json 2>> gpurun_out/r02v.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/r02v_bench_${n}_v${v}.json').read().strip().splitlines()[-1])
print($n, 'variant', $v, 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'value', d['value'])
PY
  done
done
for nf in 0.02; do
  for v in 1 0; do
    timeout 300 python bench.py --series 100000 --nan-frac $nf --steps 50 --warmup 5 --tc-variant $v --no-traffic --no-others > gpurun_out/r02v_bench_nan_v${v}.json 2>> gpurun_out/r02v.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/r02v_bench_nan_v${v}.json').read().strip().splitlines()[-1])
print('100k nan2 variant', $v, 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4))
PY
  done
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-traffic --no-others > gpurun_out/r02v_bench_default.json 2>> gpurun_out/r02v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02v_bench_default.json').read().strip().splitlines()[-1])
print('default ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
PY
tail -3 gpurun_out/r02v.err
