#!/bin/bash
# round 2, GPU run 42 (1 GPU): the whole GPU suite on the final tree with a fresh parity-error log
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02ah_pytest.log 2>&1; tail -3 gpurun_out/r02ah_pytest.log
wc -l gpurun_out/parity_errors.jsonl
