#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 1 -c 1 -o gpurun_out/prof_phi_tc $B > gpurun_out/ncu_phi.log 2>&1
echo "ncu phi exit $?"
$B > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:median_filter_tc_kernel -s 1 -c 1 -o gpurun_out/prof_median_filter $B > gpurun_out/ncu_filter.log 2>&1
echo "ncu filter exit $?"
