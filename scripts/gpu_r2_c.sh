#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --bandwidth 51.85"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_tcw_kernel -s 1 -c 1 -o gpurun_out/r2_prof_tcw_g2 $B1 > gpurun_out/ncu_tcw.log 2>&1; echo "ncu g2 exit $?"
NSVGD_TC_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_tcw_kernel -s 1 -c 1 -o gpurun_out/r2_prof_tcw_g1 $B1 > gpurun_out/ncu_tcw1.log 2>&1; echo "ncu g1 exit $?"
