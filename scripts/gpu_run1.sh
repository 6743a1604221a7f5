#!/bin/bash
# round 2, GPU run 1 (1 GPU): parity suite with measured errors, default bench line, kernel variants, small batches
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
free -g >> gpurun_out/r02a_gpu.txt; nproc >> gpurun_out/r02a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err
timeout 300 python bench.py --steps 20 --warmup 3 --tc-variant 2 --no-e2e --no-cpu-baseline --no-traffic > gpurun_out/r02a_bench_v2.json 2>> gpurun_out/r02a_bench_default.err
timeout 300 python bench.py --steps 20 --warmup 3 --tc-variant 1 --no-e2e --no-cpu-baseline --no-traffic > gpurun_out/r02a_bench_v1.json 2>> gpurun_out/r02a_bench_default.err
for s in 10000 100000; do
  timeout 300 python bench.py --steps 50 --warmup 5 --series $s --no-e2e --no-cpu-baseline --no-traffic > gpurun_out/r02a_bench_${s}.json 2>> gpurun_out/r02a_bench_default.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --nan-frac 0.02 --no-e2e --no-cpu-baseline --no-traffic > gpurun_out/r02a_bench_nan2.json 2>> gpurun_out/r02a_bench_default.err
tail -3 gpurun_out/r02a_pytest.log
cat gpurun_out/r02a_bench_default.json | head -c 3000
