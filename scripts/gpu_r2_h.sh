#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-offdiag > gpurun_out/r2_h_$tag.log 2>&1; echo -n "$tag: "; grep '^{' gpurun_out/r2_h_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); b=d['breakdown_ms_per_step']; print(d['ms_per_step'], b['median_op'], b['median_op']-b['median_tc_filter_kernel']-b['median_sample_kernel']-b['median_exact_list_kernel'], d['result']['h'])"; }
run v0 NSVGD_FSEL_VARIANT=0
run v1 NSVGD_FSEL_VARIANT=1
run v2 NSVGD_FSEL_VARIANT=2
run v3 NSVGD_FSEL_VARIANT=3
run v3m1 NSVGD_FSEL_VARIANT=3 NSVGD_FSEL_MULT=1
run v3m2 NSVGD_FSEL_VARIANT=3 NSVGD_FSEL_MULT=2
run v3m8 NSVGD_FSEL_VARIANT=3 NSVGD_FSEL_MULT=8
