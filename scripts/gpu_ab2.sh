#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -p no:cacheprovider -k "median or bandwidth" 2>&1 | tail -3
for iid in 0 1; do
  NSVGD_MEDIAN_IID_SAMPLE=$iid timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab$iid.log 2>&1; echo "iid=$iid exit $?"
  grep '^{' gpurun_out/ab$iid.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step'], d['result'])"
done
