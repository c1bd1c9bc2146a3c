#!/bin/bash
# first GPU pass: probes, SIMT/generic parity, TC parity, models, short bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run probe python -m pytest tests/test_tc_probe.py -q --tb=short -p no:cacheprovider
run parity_simt python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -k "not tc and not full_size and not far_from"
run models python -m pytest tests/test_models_gpu.py -q --tb=short -p no:cacheprovider
run parity_tc python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -k "tc or far_from"
run full python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -k "full_size"
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_small python bench.py --n 8192 --steps 3 --warmup 1 --cpu-rows 256
run bench_full python bench.py --steps 3 --warmup 1
