#!/bin/bash
# session I (1 GPU): fixed per-request overhead of the group path (tiny shards) vs shard size
set -u
O=gpurun_out
for rows in 32 9472 100000 400000 1250000; do
  timeout 200 python tools/bench_group.py --rows-per-gpu $rows --devices 0 --queries 2000 >> $O/r02i_group.jsonl 2>> $O/r02i.err
done
timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 2000 --k 1 >> $O/r02i_group.jsonl 2>> $O/r02i.err
cat $O/r02i_group.jsonl
