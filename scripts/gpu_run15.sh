#!/bin/bash
# round 2, GPU run 15 (1 GPU): ragged holdout (predict_tc MULTI) + full GPU suite + memcheck of the new path
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -25 gpurun_out/r02o_pytest.log
timeout 900 compute-sanitizer --error-exitcode 9 --tool memcheck python -m pytest tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x -k "ragged" > gpurun_out/r02o_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02o_memcheck.log
