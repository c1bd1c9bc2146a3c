#!/bin/bash
# scaling record on one 8-GPU box: N = 8, 4, 2 (graph replay, peer-memory fused all-reduces), strict timeouts
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for G in 8 4 2; do
  timeout 100 $TR --nproc-per-node $G --master-port 2953$G bench.py --gpus $G --steps 20 --warmup 3 > gpurun_out/final_n$G.log 2>&1
  echo "n$G exit $?"; grep '^{' gpurun_out/final_n$G.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step'], d['e2e']['value'], d['result'])"
done
