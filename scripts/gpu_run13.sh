#!/bin/bash
# round 2, GPU run 13 (1 GPU): racecheck after the per-thread mbarrier arrival; fit_warp early dependents A/B
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02m_racecheck.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "gap or masked or golden or graph" > gpurun_out/r02m_pytest.log 2>&1; tail -2 gpurun_out/r02m_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic --no-others"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/r02m_$name.json 2>> gpurun_out/r02m.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02m_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4))
except Exception as e: print('$name FAILED', e)
PY
}
NOP=$PWD/tests/_build/libmmf_nopdlw.so
for rep in 1 2; do
run nan2_pdl_$rep $B --nan-frac 0.02
run nan2_nopdlw_$rep env MMF_LIB=$NOP $B --nan-frac 0.02
done
run default_pdl $B
run default_nopdlw env MMF_LIB=$NOP $B
run c3_pdl $B --series 100000 --steps 50
run c3_nopdlw env MMF_LIB=$NOP $B --series 100000 --steps 50
run c2_pdl $B --series 10000 --steps 50
run c2_nopdlw env MMF_LIB=$NOP $B --series 10000 --steps 50
