#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_p2p_gpu.py -x -q --tb=short -p no:cacheprovider -k "(median or bandwidth or fullsize_fast or p2p) and not n8192" 2>&1 | tail -4
for v in 0 1 3; do
NSVGD_FSEL_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-offdiag 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); b=d['breakdown_ms_per_step']; print($v, d['ms_per_step'], b['median_op'], 'bracket', b['median_bracket_select'], 'cand', b['median_resolve_candidate_select'], 'prep', b['median_prep_kernels'], b['median_prepare'], 'phi prep/epi', b['phi_prep'], b['phi_epilogue'])"
done
