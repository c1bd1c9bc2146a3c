#!/bin/bash
# timing experiments on phi_tcw: number of j-splits (per-CTA overhead), MMAs off
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --bandwidth 51.85 > gpurun_out/r2_b_$tag.log 2>&1; echo -n "$tag: "; grep '^{' gpurun_out/r2_b_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['breakdown_ms_per_step']['phi_main_kernel'], d['breakdown_ms_per_step']['phi_op'], d['result']['ksd'])"; }
run base A=1
run split4 NSVGD_TC_NSPLIT=4
run split7 NSVGD_TC_NSPLIT=7
run split26 NSVGD_TC_NSPLIT=26
run split52 NSVGD_TC_NSPLIT=52
run nomma NSVGD_TC_DEBUG=1
run nomma_split4 NSVGD_TC_DEBUG=1 NSVGD_TC_NSPLIT=4
run g1 NSVGD_TC_GROUPS=1
run g1_split4 NSVGD_TC_GROUPS=1 NSVGD_TC_NSPLIT=4
run threepass NSVGD_TC_GRAM_EPS=0
run threepass_split4 NSVGD_TC_GRAM_EPS=0 NSVGD_TC_NSPLIT=4
