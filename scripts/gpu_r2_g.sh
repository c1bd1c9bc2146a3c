#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q --tb=short -p no:cacheprovider -k "(median or bandwidth or fullsize) and not n8192" 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-offdiag > gpurun_out/r2_g_bench.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/r2_g_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step'], d['gpu_launches'])"
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-offdiag"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_g.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
