#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference_golden.py -x -q --tb=short -p no:cacheprovider -k "not n8192" 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-offdiag 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); b=d['breakdown_ms_per_step']; print(d['ms_per_step'], d['e2e']['value']/d['value']); [print(' ',k,v) for k,v in b.items()]"
