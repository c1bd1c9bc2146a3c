#!/bin/bash
# multi-GPU sanity + scaling: torchrun bench on N GPUs (N = $1)
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "n1 exit $?"; tail -1 gpurun_out/bench_n1.log | cut -c1-700
for G in 2 4 8; do
  if [ $G -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 5 --warmup 2 > gpurun_out/bench_n$G.log 2>&1; echo "n$G exit $?"; grep '^{' gpurun_out/bench_n$G.log | cut -c1-900; tail -3 gpurun_out/bench_n$G.log | cut -c1-300
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_sharded.py > gpurun_out/check_sharded.log 2>&1; echo "check exit $?"; tail -5 gpurun_out/check_sharded.log
