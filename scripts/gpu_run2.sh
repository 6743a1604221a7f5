#!/bin/bash
# round 2, GPU run 2 (2 GPUs): fused multi-GPU parity vs oracle + 2-GPU bench lines (weak / strong, p2p / multicast-bulk)
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
nvidia-smi topo -m > gpurun_out/r02b_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "symmetric or captured or mostly_missing or full_series or broadcast" > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r02b_bench_2gpu_p2p.json 2> gpurun_out/r02b_bench.err
timeout 400 $TR --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --gather multicast-bulk > gpurun_out/r02b_bench_2gpu_mcbulk.json 2>> gpurun_out/r02b_bench.err
timeout 400 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --scaling strong > gpurun_out/r02b_bench_2gpu_strong.json 2>> gpurun_out/r02b_bench.err
timeout 400 $TR --master-port 29704 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --tc-variant 1 > gpurun_out/r02b_bench_2gpu_p2p_v1.json 2>> gpurun_out/r02b_bench.err
tail -5 gpurun_out/r02b_pytest.log
for f in gpurun_out/r02b_bench_2gpu_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['config']['gather'], d['config']['gather_max_abs_diff_vs_nccl'], d.get('shard_only',{}).get('ms_per_step'), d.get('shard_only',{}).get('nvlink'))
"; done
tail -5 gpurun_out/r02b_bench.err
