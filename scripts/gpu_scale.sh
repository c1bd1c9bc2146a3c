#!/bin/bash
# scaling run on N GPUs (N = $1): sharded check, then bench at G = 2,4,8 <= N eager and graphed
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpus_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $TR --nproc-per-node 2 --master-port 29512 scripts/check_sharded.py > gpurun_out/check_sharded.log 2>&1; echo "check exit $?"; grep "rank 0" gpurun_out/check_sharded.log | tail -5
for G in 2 4 8; do
  if [ $G -le $N ]; then
    for mode in off on; do
      timeout 150 $TR --nproc-per-node $G --master-port 2951$G bench.py --gpus $G --steps 20 --warmup 3 --graph $mode --no-e2e > gpurun_out/scale_n${G}_graph_$mode.log 2>&1
      echo "n$G graph=$mode exit $?"; grep '^{' gpurun_out/scale_n${G}_graph_$mode.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown_ms_per_step'], d['wall_s'], d['gpu_launches'])"
    done
  fi
done
