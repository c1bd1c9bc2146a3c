#!/bin/bash
# scaling run: bash scripts/gpu_scale.sh "<list of G>" [check]   (bench eager and graphed at each G; strict timeouts)
GS=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ -n "$2" ]; then
  timeout 150 $TR --nproc-per-node 2 --master-port 29512 scripts/check_sharded.py > gpurun_out/check_sharded.log 2>&1; echo "check exit $?"; grep "rank 0" gpurun_out/check_sharded.log | tail -5
fi
for G in $GS; do
  for p2p in 1 0; do for mode in on off; do
    NSVGD_P2P=$p2p timeout 120 $TR --nproc-per-node $G --master-port 2951$G bench.py --gpus $G --steps 10 --warmup 3 --graph $mode --no-e2e > gpurun_out/scale_n${G}_p2p${p2p}_graph_$mode.log 2>&1
    echo "n$G p2p=$p2p graph=$mode exit $?"; grep '^{' gpurun_out/scale_n${G}_p2p${p2p}_graph_$mode.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['small_allreduces'][:12], d['breakdown_ms_per_step']['phi_op'], d['wall_s'], d['gpu_launches'], d['result'])"
  done; done
done
