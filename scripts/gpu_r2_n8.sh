#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_p2p_gpu.py -x -q --tb=short -p no:cacheprovider 2>&1 | tail -3
for N in 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 30 --warmup 5 --no-offdiag > gpurun_out/r2_n$N.log 2>&1; echo "N=$N exit $?"
grep "^{" gpurun_out/r2_n$N.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value']/d['value'], d['gpu_launches']); [print(' ',k,v) for k,v in d['breakdown_ms_per_step'].items()]"
done
