#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python scripts/profile_shard.py 8 5 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_shard8.csv python scripts/profile_shard.py 8 1 > gpurun_out/ncu_shard.log 2>&1; echo "ncu exit $?"
