#!/bin/bash
# round 2, GPU run 24 (1 GPU): gap segments filed by chunk parity (position-independent forecasts), balanced launches for
# sub-wave batches only; whole GPU suite; small-batch and gappy lines; default line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02w_pytest.log 2>&1; tail -5 gpurun_out/r02w_pytest.log
python scripts/dbg_bal.py 2>&1 | grep differing
for n in 10000 100000; do
  timeout 300 python bench.py --series $n --steps 50 --warmup 5 --no-traffic --no-others > gpurun_out/r02w_bench_${n}.json 2>> gpurun_out/r02w.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02w_bench_${n}.json').read().strip().splitlines()[-1])
print($n, 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'value', d['value'])
PY
done
timeout 300 python bench.py --nan-frac 0.02 --steps 30 --warmup 5 --no-traffic --no-others > gpurun_out/r02w_bench_nan2.json 2>> gpurun_out/r02w.err
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r02w_bench_default.json 2>> gpurun_out/r02w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02w_bench_nan2.json').read().strip().splitlines()[-1])
print('nan2 ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
d=json.loads(open('gpurun_out/r02w_bench_default.json').read().strip().splitlines()[-1])
print('default ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'others', d.get('other_configs'))
PY
tail -3 gpurun_out/r02w.err
