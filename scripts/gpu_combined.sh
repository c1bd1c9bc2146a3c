#!/bin/bash
# combined-W phi kernel: parity, A/B bench, ncu capture
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-1500; }
run parity python -m pytest tests/test_gpu_parity.py tests/test_models_gpu.py tests/test_gpu_reference_golden.py -x -q --tb=short -p no:cacheprovider
run bench_combined python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
NSVGD_TC_COMBINE=0 run bench_separate python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 1 -c 1 -o gpurun_out/prof_phi_tc_combined $B > gpurun_out/ncu_phi.log 2>&1
echo "ncu phi exit $?"
