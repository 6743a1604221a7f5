#!/bin/bash
# round 2, GPU run 17 (1 GPU): hybrid transport sweep (every Nth chunk float32), widen kernel rewrite, full suite
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02q_pytest.log
tail -6 gpurun_out/r02q_pytest.log
timeout 300 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:widen_kernel -c 2 --csv --print-units base --log-file gpurun_out/r02q_widen.csv python scripts/ncu_scenarios.py widen 500000 > /dev/null 2>> gpurun_out/r02q.err
grep -v "^==" gpurun_out/r02q_widen.csv | cut -d, -f5,13-15 | tail -6
cat > /tmp/sweep.py <<'PY'
import sys, time, os, numpy as np
sys.path.insert(0,'.')
import mmf, torch
n,t,h=1000000,1095,28
y,start=mmf.synth.daily_store_item_demand_torch(n,t,seed=1)
mmf.bind_to_gpu_numa(0)
yh=mmf.alloc_packed(n,t); yh[...]=y.cpu().numpy(); oh=mmf.pinned_empty((n,h))
ref=None
for rep in range(2):
  for de in [0, 8, 6, 5, 4, 3, 2]:
    os.environ["MMF_HOST_DIRECT_EVERY"]=str(de)
    eng=mmf.ForecastEngine(host_narrow="on")
    _,ps,npred=eng.plan_calendar(start,t,"D",h,"future")
    for _ in range(2): eng.fit_forecast(yh,ps,npred,out=oh)
    t0=time.perf_counter()
    for _ in range(5): eng.fit_forecast(yh,ps,npred,out=oh)
    dt=(time.perf_counter()-t0)/5
    if ref is None: ref=oh.copy()
    print("direct_every",de,"ms/step",round(dt*1e3,2),"series/s",round(n/dt/1e6,2),"M", "equal", bool(np.array_equal(ref,oh)),flush=True)
    eng.close()
eng=mmf.ForecastEngine(host_narrow="off")
_,ps,npred=eng.plan_calendar(start,t,"D",h,"future")
for _ in range(2): eng.fit_forecast(yh,ps,npred,out=oh)
t0=time.perf_counter()
for _ in range(5): eng.fit_forecast(yh,ps,npred,out=oh)
dt=(time.perf_counter()-t0)/5
print("narrow off ms/step",round(dt*1e3,2),"series/s",round(n/dt/1e6,2),"M", "equal", bool(np.array_equal(ref,oh)),flush=True)
PY
python /tmp/sweep.py
