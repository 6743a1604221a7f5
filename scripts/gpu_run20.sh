#!/bin/bash
# round 2, GPU run 20 (1 GPU): DataFrame boundary after the single-group fast path; frames tests on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "forecast_groups or spark or apply_in_pandas or packer or selection or exog" > gpurun_out/r02t_pytest.log 2>&1; tail -3 gpurun_out/r02t_pytest.log
timeout 900 python scripts/bench_frames.py 20000 157 > gpurun_out/r02t_bench_frames_20k.json 2> gpurun_out/r02t.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02t_bench_frames_20k.json').read().strip().splitlines()[-1])
for k,v in d.items(): print(k, v)
PY
tail -3 gpurun_out/r02t.err
