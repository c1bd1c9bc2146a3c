#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --bandwidth 51.85 > gpurun_out/r2_e_$tag.log 2>&1; echo -n "$tag: "; grep '^{' gpurun_out/r2_e_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['breakdown_ms_per_step']['phi_main_kernel'], d['breakdown_ms_per_step']['phi_op'], d['result']['ksd'])"; }
run base A=1
run noV1 NSVGD_TC_DEBUG=4
run noV NSVGD_TC_DEBUG=12
run nomma_noV NSVGD_TC_DEBUG=13
run nomma NSVGD_TC_DEBUG=1
B1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --bandwidth 51.85"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_tcw_kernel -s 1 -c 1 -o gpurun_out/r2_prof_tcw_2iss $B1 > gpurun_out/ncu_tcw.log 2>&1; echo "ncu exit $?"
