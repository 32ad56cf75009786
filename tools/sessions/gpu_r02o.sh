#!/bin/bash
# session O (1 GPU): L2-resident slice of the shard (evict_last tiles) -- sweep of the slice size (tuning build)
set -u
O=gpurun_out
T=multimodal-image-similarity-search_b200/libvecsearch_b200_tuning.so
timeout 600 python -m pytest tests/test_gpu_scan_parity.py tests/test_gpu_exchange.py -x -q -m gpu > $O/r02o_tests.log 2>&1; echo "tests rc=$?" >> $O/r02o_tests.log
for mb in 0 32 48 64 80 96 112; do
  VS_SCAN_KEEP_MB=$mb VS_LIB_PATH=$T timeout 200 python tools/bench_scan.py --rows 1250000,10000000 --queries 32 --iters 8 >> $O/r02o_scan.jsonl 2>> $O/r02o.err
done
for mb in 0 48 64 96; do
  echo "## VS_SCAN_KEEP_MB=$mb" >> $O/r02o_group.jsonl
  VS_SCAN_KEEP_MB=$mb VS_LIB_PATH=$T timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 2000 >> $O/r02o_group.jsonl 2>> $O/r02o.err
done
VS_SCAN_KEEP_MB=0 VS_LIB_PATH=$T timeout 200 python tools/bench_scan.py --rows 1000000 --dtype f32 --queries 32 --iters 8 >> $O/r02o_scan.jsonl 2>> $O/r02o.err
VS_SCAN_KEEP_MB=64 VS_LIB_PATH=$T timeout 200 python tools/bench_scan.py --rows 1000000 --dtype f32 --queries 32 --iters 8 >> $O/r02o_scan.jsonl 2>> $O/r02o.err
tail -3 $O/r02o_tests.log; cat $O/r02o_scan.jsonl $O/r02o_group.jsonl; tail -5 $O/r02o.err
