#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
: > gpurun_out/shard_probe.log
for t in 1 2 4 8; do timeout 300 python tools/shard_probe.py $t 32 >> gpurun_out/shard_probe.log 2>&1; done
for t in 2 8; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_shard_tp$t.csv python tools/shard_probe.py $t 1 > /dev/null 2>&1
done
cat gpurun_out/shard_probe.log
python tools/agg_launches.py gpurun_out/launches_shard_tp2.csv
python tools/agg_launches.py gpurun_out/launches_shard_tp8.csv
