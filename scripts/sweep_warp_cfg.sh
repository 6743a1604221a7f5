#!/bin/bash
# usage: sweep.sh "<S> <U> <W>" ...   (runs on the GPU box; rebuilds fit_warp.o per config)
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  touch dss-ml-at-scale_b200/csrc/fit_warp.cu
  make -C dss-ml-at-scale_b200/csrc ../libmmf.so EXTRA="-DMMF_WARP_S=$1 -DMMF_WARP_U=$2 -DMMF_WARP_WARPS=$3" > /dev/null 2>&1
  for extra in "--kernel warp" "--nan-frac 0.02 --kernel warp"; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $extra 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('S=$1 U=$2 W=$3', '$extra', round(d['ms_per_step'],3), 'ms', round(d['roofline']['frac'],3))"
  done
done
