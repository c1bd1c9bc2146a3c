#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -p no:cacheprovider -k "phi_paths and tc or full_size or far_from" 2>&1 | tail -3
for pp in 1 0; do
  NSVGD_TC_PINGPONG=$pp timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/pp$pp.log 2>&1
  echo -n "pingpong=$pp: "; grep '^{' gpurun_out/pp$pp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step']['phi_main_kernel'], d['breakdown_ms_per_step']['phi_op'], d['result'])"
done
for dbg in 1 2 3; do
  NSVGD_TC_DEBUG=$dbg timeout 200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --bandwidth 50 > gpurun_out/dbg$dbg.log 2>&1
  echo -n "pingpong dbg=$dbg: "; grep '^{' gpurun_out/dbg$dbg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step']['phi_main_kernel'])"
done
