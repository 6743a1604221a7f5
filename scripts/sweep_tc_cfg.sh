#!/bin/bash
# usage: sweep_tc_cfg.sh "<STAGES> <ASLOTS>" ...   (runs on the GPU box; rebuilds fit_tc.o per config)
for cfg in "$@"; do
  set -- $cfg
  touch dss-ml-at-scale_b200/csrc/fit_tc.cu
  make -C dss-ml-at-scale_b200/csrc ../libmmf.so EXTRA="-DMMF_TC_STAGES=$1 -DMMF_TC_ASLOTS=$2 $EXTRA_DEFS" > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('STAGES=$1 ASLOTS=$2', round(d['roofline']['kernel_ms'],4), 'ms', round(d['roofline']['achieved'],1), 'GB/s', round(d['roofline']['frac'],3))"
done
