#!/bin/bash
# last sanity of the round: all GPU tests, smoke, both bench arms, SIMT-path profile
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-300; }
run pytest_gpu python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
run bench_full python bench.py
run simt python scripts/profile_simt.py
timeout 300 ncu --set full --clock-control none -k regex:phi_smalld -s 1 -c 1 -o gpurun_out/prof_phi_smalld python scripts/profile_simt.py > gpurun_out/ncu_simt.log 2>&1; echo "ncu simt exit $?"
