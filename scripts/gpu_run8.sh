#!/bin/bash
# round 2, GPU run 8 (1 GPU): host narrowing -- sub-chunk size / store mode / thread sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k "narrow or integer" > gpurun_out/r02h_pytest.log 2>&1
tail -3 gpurun_out/r02h_pytest.log
cat > /tmp/sweep.py <<'PY'
import sys, time, os, numpy as np
sys.path.insert(0,'.')
import mmf, torch
n,t,h=1000000,1095,28
y,start=mmf.synth.daily_store_item_demand_torch(n,t,seed=1)
mmf.bind_to_gpu_numa(0)
yh=mmf.alloc_packed(n,t); yh[...]=y.cpu().numpy(); oh=mmf.pinned_empty((n,h))
for th, sub, st in [(16,4096,0),(16,4096,1),(16,2048,0),(16,8192,0),(16,1024,0),(16,32768,1),(16,32768,0),(8,4096,0),(24,4096,0),(32,4096,0),(12,4096,0)]:
    os.environ["MMF_HOST_SUB_ROWS"]=str(sub); os.environ["MMF_HOST_STREAM_STORES"]=str(st)
    eng=mmf.ForecastEngine(host_narrow="on", host_threads=th)
    _,ps,npred=eng.plan_calendar(start,t,"D",h,"future")
    for _ in range(2): eng.fit_forecast(yh,ps,npred,out=oh)
    t0=time.perf_counter()
    for _ in range(5): eng.fit_forecast(yh,ps,npred,out=oh)
    dt=(time.perf_counter()-t0)/5
    print("threads",th,"sub_rows",sub,"stream_stores",st,"ms/step",round(dt*1e3,2),"series/s",round(n/dt/1e6,2),"M",flush=True)
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
