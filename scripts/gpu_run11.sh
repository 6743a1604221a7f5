#!/bin/bash
# round 2, GPU run 11 (8 GPUs): fused multi-GPU table vs the oracle on 8 ranks; BASELINE configs[4] as worded (10 M x 365 over 8 GPUs,
# device-resident and spilled to pinned host memory); configs[3] e2e at 8 ranks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -k "symmetric" > gpurun_out/r02k_pytest.log 2>&1
tail -3 gpurun_out/r02k_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; np=$2; shift 2
  timeout 600 $TR --nproc-per-node $np --master-port $((29800 + RANDOM % 100)) bench.py --gpus $np "$@" > gpurun_out/r02k_${name}.json 2>> gpurun_out/r02k.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02k_${name}.json').read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print('${name}', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e9,3),'G', 'fit_only', round((d.get('shard_only') or {}).get('ms_per_step',0),4), 'e2e', round(e.get('value',0)/1e6,2),'M', e.get('ms_per_step'), e.get('h2d_bytes_per_step'))
except Exception as e:
    print('${name}', 'FAILED', e)
PY
}
run cfg5_8gpu_strong 8 --series 10000000 --t 365 --scaling strong --steps 20 --warmup 3
run cfg4_8gpu_weak_e2e 8 --steps 10 --warmup 3
run cfg5_4gpu_strong 4 --series 10000000 --t 365 --scaling strong --steps 20 --warmup 3
grep -v "^\*\|OMP\|^$" gpurun_out/r02k.err | tail -5
