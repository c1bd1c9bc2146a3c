#!/bin/bash
# 8-GPU record: graph replay, peer-memory vs NCCL small all-reduces (strict timeouts)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for p2p in 1 0; do
  NSVGD_P2P=$p2p timeout 100 $TR --nproc-per-node 8 --master-port 2952$p2p bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/scale_n8_p2p${p2p}.log 2>&1
  echo "n8 p2p=$p2p exit $?"; grep '^{' gpurun_out/scale_n8_p2p${p2p}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['small_allreduces'][:12], d['breakdown_ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['result'])"
done
