#!/bin/bash
# round 2: new COMBINED-form kernel (phi_tcw): parity subset, full-size parity, bench A/B against the round-1 kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py tests/test_models_gpu.py -x -q --tb=short -p no:cacheprovider -k "not fast_median" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q --tb=short -p no:cacheprovider --durations=8 2>&1 | tail -40
for v in v2 v1; do
NSVGD_TC_KERNEL=$v timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_a_$v.log 2>&1; echo "bench $v exit $?"
grep '^{' gpurun_out/r2_a_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step']['phi_main_kernel'], d['breakdown_ms_per_step']['phi_op'], d['result'])"
NSVGD_TC_KERNEL=$v timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --median-mode offdiag > gpurun_out/r2_a_${v}_off.log 2>&1; echo "bench $v offdiag exit $?"
grep '^{' gpurun_out/r2_a_${v}_off.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step']['phi_main_kernel'], d['breakdown_ms_per_step']['phi_op'], d['result'])"
done
tail -5 gpurun_out/r2_a_v2.log | cut -c1-600
