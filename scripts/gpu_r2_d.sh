#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -p no:cacheprovider -k "tc or golden or step" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q --tb=short -p no:cacheprovider 2>&1 | tail -5
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --bandwidth 51.85 > gpurun_out/r2_d_$tag.log 2>&1; echo -n "$tag: "; grep '^{' gpurun_out/r2_d_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['breakdown_ms_per_step']['phi_main_kernel'], d['breakdown_ms_per_step']['phi_op'], d['result']['ksd'])"; }
run base A=1
run g1 NSVGD_TC_GROUPS=1
run split13 NSVGD_TC_NSPLIT=13
run split1 NSVGD_TC_NSPLIT=1
run threepass NSVGD_TC_GRAM_EPS=0
run threepass_g1 NSVGD_TC_GRAM_EPS=0 NSVGD_TC_GROUPS=1
