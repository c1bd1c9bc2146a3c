#!/bin/bash
# round 2, call 1: state of the shipped kernels before any change (ncu of the ONE-pass phi kernel, filter, launch list)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench0.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/r2_bench0.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step'], d['roofline']['avg_kernel_ms'], d['clocks'])"
timeout 120 python bench.py --steps 20 --warmup 5 --median-mode offdiag --no-e2e --no-cpu-baseline > gpurun_out/r2_bench0_offdiag.log 2>&1; echo "bench offdiag exit $?"
grep '^{' gpurun_out/r2_bench0_offdiag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step'], d['roofline']['avg_kernel_ms'], d['result'])"
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches0.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
B1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_tc_kernel -s 1 -c 1 -o gpurun_out/r2_prof_phi_tc0 $B1 > gpurun_out/ncu_phi.log 2>&1; echo "ncu phi exit $?"
