#!/bin/bash
# round 2, GPU run 27 (1 GPU): compute-sanitizer on the final tree -- gap segments by chunk parity, the balanced experiment variant
mkdir -p gpurun_out
timeout 900 compute-sanitizer --error-exitcode 9 --tool memcheck python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k "balanced and not 10000 and not 20011" > gpurun_out/r02z_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02z_memcheck.log
cat > /tmp/rc.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mmf
for variant in (1, 3):
    yd, start = mmf.synth.daily_store_item_demand_torch(700, 400, seed=3, nan_frac=0.02)
    eng = mmf.ForecastEngine(kernel="tc", tc_variant=variant)
    r = mmf.forecast_packed(yd, start, "D", 28, "future", engine=eng, want_status=True)
    torch.cuda.synchronize()
    print(variant, float(r["pred"].nan_to_num().abs().sum()), int((r["status"] == 0).sum()))
    eng.close()
PY
timeout 900 compute-sanitizer --error-exitcode 9 --tool racecheck python /tmp/rc.py > gpurun_out/r02z_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r02z_racecheck.log
timeout 600 compute-sanitizer --error-exitcode 9 --tool synccheck python /tmp/rc.py > gpurun_out/r02z_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r02z_synccheck.log
