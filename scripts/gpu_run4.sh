#!/bin/bash
# round 2, GPU run 4 (8 GPUs): weak / strong scaling lines at N = 8 and 4, gather variants
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --steps 20 --warmup 3 --no-e2e"
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 300 $TR --nproc-per-node $np --master-port $((29800 + RANDOM % 100)) $B --gpus $np "$@" > gpurun_out/r02d_${name}.json 2>> gpurun_out/r02d_bench.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02d_${name}.json').read().strip().splitlines()[-1])
    so=d.get('shard_only') or {}
    print('${name}', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e9,3),'G', 'fit_ms', round(d['roofline']['kernel_ms'],4), 'shard_only', round(so.get('ms_per_step',0),4), 'diff', d['config']['gather_max_abs_diff_vs_nccl'], 'nvlink', so.get('nvlink'))
except Exception as e:
    print('${name}', 'FAILED', e)
PY
}
run 8_p2p_auto 8
run 8_p2p_v1 8 --tc-variant 1
run 8_mcbulk 8 --gather multicast-bulk
run 4_p2p_auto 4
run 4_p2p_v1 4 --tc-variant 1
run 8_strong 8 --scaling strong
run 4_strong 4 --scaling strong
run 8_nccl 8 --gather nccl
grep -v "^\*\|OMP\|^$" gpurun_out/r02d_bench.err | tail -5
