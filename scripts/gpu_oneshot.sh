#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 40 python -m pytest tests/test_gpu_parity.py -x -q --tb=line -p no:cacheprovider -k "phi_paths and tc" 2>&1 | tail -3
timeout 25 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab.log 2>&1; echo "exit $?"
grep '^{' gpurun_out/ab.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step']['phi_main_kernel'], d['breakdown_ms_per_step']['phi_op'], d['result'])"
