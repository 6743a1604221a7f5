#!/bin/bash
# round 2, GPU run 21 (2 GPUs): the driver's multi-GPU invocation on the final tree (default line + reference arm under torchrun)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02u_bench_2gpu.json 2> gpurun_out/r02u.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02u_bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu', d['ms_per_step'], d['value'], d['scaling'], d['config']['gather'], d['config']['gather_max_abs_diff_vs_nccl'], 'shard_only', d['shard_only']['ms_per_step'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e'].get('transport','')[:60])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
timeout 600 $TR --master-port 29722 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02u_ref_2gpu.json 2>> gpurun_out/r02u.err; cut -c1-160 gpurun_out/r02u_ref_2gpu.json
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -k symmetric > gpurun_out/r02u_pytest.log 2>&1; tail -2 gpurun_out/r02u_pytest.log
grep -v "^\*\|OMP\|^$" gpurun_out/r02u.err | tail -4
