#!/bin/bash
# session U (1 GPU): same-box A/B of (HEAD) vs (TMA producer starts before the query staging) vs (+ L2-resident slice)
set -u
O=gpurun_out
P=multimodal-image-similarity-search_b200
timeout 900 python -m pytest tests/test_gpu_scan_parity.py tests/test_gpu_exchange.py tests/test_gpu_group.py -x -q -m gpu > $O/r02u_tests.log 2>&1; echo "tests rc=$?" >> $O/r02u_tests.log
g() { echo "## $1" >> $O/r02u_group.jsonl; shift; env "$@" timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 2000 >> $O/r02u_group.jsonl 2>> $O/r02u.err; }
for rep in 1 2; do
  g head VS_LIB_PATH=$P/libvecsearch_b200_head.so
  g early_producer_only VS_LIB_PATH=$P/libvecsearch_b200_tuning.so VS_SCAN_KEEP_MB=0
  g product_both VS_X=1
done
s() { echo "## $1" >> $O/r02u_scan.jsonl; shift; env "$@" timeout 200 python tools/bench_scan.py --rows 1250000,10000000 --queries 32 --iters 8 >> $O/r02u_scan.jsonl 2>> $O/r02u.err; }
s head VS_LIB_PATH=$P/libvecsearch_b200_head.so
s product VS_X=1
s head VS_LIB_PATH=$P/libvecsearch_b200_head.so
s product VS_X=1
tail -3 $O/r02u_tests.log; cat $O/r02u_group.jsonl | cut -c1-420; cat $O/r02u_scan.jsonl; tail -5 $O/r02u.err
