#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_p2p_gpu.py tests/test_gpu_fullsize.py tests/test_sharded_gpu.py -x -q --tb=short -p no:cacheprovider -k "median or bandwidth or p2p or fullsize or sharded" 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_f_bench.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/r2_f_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step'], d['gpu_launches'], d['offdiag_mode']['ms_per_step'], d['offdiag_mode']['median_op_ms'])"
