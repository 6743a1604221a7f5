#!/bin/bash
# round 2, GPU run 38 (1 GPU): device packer against pandas asfreq on random frames (hypothesis); packer tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "packer" > gpurun_out/r02af_pytest.log 2>&1; tail -15 gpurun_out/r02af_pytest.log
