#!/bin/bash
# full GPU verification: every gpu test, smoke, the two bench arms (what the driver runs at round end)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-600; }
run pytest_gpu python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider --durations=8
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run bench_full python bench.py
