#!/bin/bash
# round 2, GPU run 44 (2 GPUs): the driver's torchrun invocation of the default line on the final bench.py
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02aj_bench_2gpu.json 2> gpurun_out/r02aj.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02aj_bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu', d['ms_per_step'], d['value'], d['scaling'], d['config']['gather'], d['config']['gather_max_abs_diff_vs_nccl'], 'shard_only', d['shard_only']['ms_per_step'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['host_narrow'])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
grep -v "^\*\|OMP\|^$" gpurun_out/r02aj.err | tail -4
