#!/bin/bash
# round 2, GPU run 7 (1 GPU): host narrowing thread sweep after NT stores / prefetch / pinned workers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k "narrow or integer or config5" > gpurun_out/r02g_pytest.log 2>&1
tail -3 gpurun_out/r02g_pytest.log
for th in 8 16 24 32 48; do
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
print("host_threads", $th, "ms/step", round(dt*1e3,2), "series/s", round(n/dt/1e6,2), "M", flush=True)
PY
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-traffic > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['max_abs_diff_vs_device_path'], d['e2e']['uint16_ingest'])
PY
