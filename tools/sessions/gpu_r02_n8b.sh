#!/bin/bash
# short 8-GPU session: per-request latency of the single-process group after removing the per-launch attribute call
set -u
O=gpurun_out
timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --queries 3000 >> $O/r02_n8b_group.jsonl 2>> $O/r02_n8b.err
timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --queries 3000 --devices 0,1,2,3 >> $O/r02_n8b_group.jsonl 2>> $O/r02_n8b.err
timeout 200 python tools/bench_group.py --rows-per-gpu 32 --queries 3000 >> $O/r02_n8b_group.jsonl 2>> $O/r02_n8b.err
cat $O/r02_n8b_group.jsonl
