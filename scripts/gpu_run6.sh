#!/bin/bash
# round 2, GPU run 6 (1 GPU): host narrowing, faster select/widen, default bench line
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" 
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -6 gpurun_out/r02f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('e2e', json.dumps(d['e2e'])[:1500])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
for th in 4 8 16 32; do
python - <<PY
import sys, time, numpy as np
sys.path.insert(0,'.')
import mmf, torch
n,t,h=1000000,1095,28
y,start=mmf.synth.daily_store_item_demand_torch(n,t,seed=1)
mmf.bind_to_gpu_numa(0)
eng=mmf.ForecastEngine(host_narrow="on", host_threads=$th)
_,ps,npred=eng.plan_calendar(start,t,"D",h,"future")
yh=mmf.alloc_packed(n,t); yh[...]=y.cpu().numpy(); oh=mmf.pinned_empty((n,h))
for _ in range(2): eng.fit_forecast(yh,ps,npred,out=oh)
t0=time.perf_counter()
for _ in range(5): eng.fit_forecast(yh,ps,npred,out=oh)
dt=(time.perf_counter()-t0)/5
print("host_threads", $th, "ms/step", round(dt*1e3,2), "series/s", round(n/dt/1e6,2), "M")
PY
done
timeout 300 python scripts/bench_packer.py 100000 365 > gpurun_out/r02f_packer_100k_x365.json 2>> gpurun_out/r02f.err; cat gpurun_out/r02f_packer_100k_x365.json
tail -3 gpurun_out/r02f.err
