#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-1800; }
run bench_full python bench.py --steps 5 --warmup 2
run all_gpu_tests python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 1 -c 1 -o gpurun_out/prof_phi_tc $B > gpurun_out/ncu_phi.log 2>&1
echo "ncu phi exit $?"
$B > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:median_filter_tc_kernel -s 1 -c 1 -o gpurun_out/prof_median_filter $B > gpurun_out/ncu_filter.log 2>&1
echo "ncu filter exit $?"
