#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -6 gpurun_out/$name.log; }
run median_fast python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -k "fast_median or bandwidth"
run parity_tc python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -k "tc or far_from or full_size"
run bench_full python bench.py --steps 5 --warmup 2
run bench_fixedh python bench.py --steps 5 --warmup 2 --bandwidth 6.0 --no-cpu-baseline
