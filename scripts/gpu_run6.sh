#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -6 gpurun_out/$name.log | cut -c1-600; }
run tune_phi python scripts/tune_phi.py
run parity_tc python -m pytest tests/test_gpu_parity.py tests/test_tc_probe.py -q --tb=short -p no:cacheprovider -k "tc or far_from or full_size or probe or staged"
run bench_full python bench.py --steps 5 --warmup 2 --no-cpu-baseline
