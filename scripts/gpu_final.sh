#!/bin/bash
# round-end record on one B200: all GPU tests, smoke, both bench arms, ncu launch list + full captures
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-400; }
run pytest_gpu python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider --durations=5
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run bench_full python bench.py
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
B1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 1 -c 1 -o gpurun_out/prof_phi_tc_final $B1 > gpurun_out/ncu_phi.log 2>&1; echo "ncu phi exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:median_filter_tc_kernel -s 1 -c 1 -o gpurun_out/prof_filter_final $B1 > gpurun_out/ncu_filter.log 2>&1; echo "ncu filter exit $?"
timeout 600 ncu --set full --clock-control none -k regex:exact_list_kernel -s 1 -c 1 -o gpurun_out/prof_exact_final $B1 > gpurun_out/ncu_exact.log 2>&1; echo "ncu exact exit $?"
