#!/bin/bash
# round 2, GPU run 14 (1 GPU): narrowing pool pinned / unpinned; default bench line on the final build
mkdir -p gpurun_out
cat > /tmp/sweep.py <<'PY'
import sys, time, os, numpy as np
sys.path.insert(0,'.')
import mmf, torch
n,t,h=1000000,1095,28
y,start=mmf.synth.daily_store_item_demand_torch(n,t,seed=1)
mmf.bind_to_gpu_numa(0)
yh=mmf.alloc_packed(n,t); yh[...]=y.cpu().numpy(); oh=mmf.pinned_empty((n,h))
for rep in range(2):
  for pin, sub, th in [(1,8192,16),(0,8192,16),(1,32768,16),(0,32768,16),(1,8192,12),(1,8192,20)]:
    os.environ["MMF_HOST_SUB_ROWS"]=str(sub); os.environ["MMF_HOST_PIN"]=str(pin)
    eng=mmf.ForecastEngine(host_narrow="on", host_threads=th)
    _,ps,npred=eng.plan_calendar(start,t,"D",h,"future")
    for _ in range(2): eng.fit_forecast(yh,ps,npred,out=oh)
    t0=time.perf_counter()
    for _ in range(5): eng.fit_forecast(yh,ps,npred,out=oh)
    dt=(time.perf_counter()-t0)/5
    print("pin",pin,"sub_rows",sub,"threads",th,"ms/step",round(dt*1e3,2),"series/s",round(n/dt/1e6,2),"M",flush=True)
    eng.close()
eng=mmf.ForecastEngine(host_narrow="off")
_,ps,npred=eng.plan_calendar(start,t,"D",h,"future")
for _ in range(2): eng.fit_forecast(yh,ps,npred,out=oh)
t0=time.perf_counter()
for _ in range(5): eng.fit_forecast(yh,ps,npred,out=oh)
dt=(time.perf_counter()-t0)/5
print("narrow off ms/step",round(dt*1e3,2),"series/s",round(n/dt/1e6,2),"M",flush=True)
PY
python /tmp/sweep.py
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02n_bench_default.json 2> gpurun_out/r02n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02n_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['max_abs_diff_vs_device_path'], d['e2e']['uint16_ingest']['value'])
print('cpu', d['cpu_baseline']['value'])
for k,v in (d.get('other_configs') or {}).items(): print(k, v)
PY
