#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for pp in 0 1; do for dbg in 4 5 6 7; do
  NSVGD_TC_PINGPONG=$pp NSVGD_TC_DEBUG=$dbg timeout 200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --bandwidth 50 > gpurun_out/dbg$dbg.log 2>&1
  echo -n "pp=$pp dbg=$dbg: "; grep '^{' gpurun_out/dbg$dbg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step']['phi_main_kernel'])"
done; done
