#!/bin/bash
# round 2, GPU run 3 (2 GPUs): new GPU tests (integer ingest, packer nulls/duplicates, spark harness) + 2-GPU bench lines
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "integer or null_keys or spark or apply_in_pandas or captured or mostly_missing or packer" > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r02c_bench_2gpu_p2p.json 2> gpurun_out/r02c_bench.err
timeout 400 $TR --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --gather multicast-bulk > gpurun_out/r02c_bench_2gpu_mcbulk.json 2>> gpurun_out/r02c_bench.err
timeout 400 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --scaling strong > gpurun_out/r02c_bench_2gpu_strong.json 2>> gpurun_out/r02c_bench.err
timeout 400 $TR --master-port 29704 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --tc-variant 1 > gpurun_out/r02c_bench_2gpu_p2p_v1.json 2>> gpurun_out/r02c_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-traffic > gpurun_out/r02c_bench_1gpu_e2e.json 2>> gpurun_out/r02c_bench.err
tail -5 gpurun_out/r02c_pytest.log
for f in gpurun_out/r02c_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['config']['gather'], d['config']['gather_max_abs_diff_vs_nccl'], d.get('shard_only',{}).get('ms_per_step'), d.get('shard_only',{}).get('nvlink'), d.get('e2e'))
"; done
grep -v "^\*\|OMP\|^$" gpurun_out/r02c_bench.err | tail -5
