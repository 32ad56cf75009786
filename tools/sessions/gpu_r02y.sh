#!/bin/bash
# session Y (1 GPU): cursor-based selection on the query tail -- parity tests, phase marks, same-box A/B against HEAD
set -u
O=gpurun_out
P=multimodal-image-similarity-search_b200
timeout 1200 python -m pytest tests/test_gpu_scan_parity.py tests/test_gpu_exchange.py tests/test_gpu_group.py tests/test_gpu_collection.py -x -q -m gpu > $O/r02y_tests.log 2>&1; echo "tests rc=$?" >> $O/r02y_tests.log
tail -3 $O/r02y_tests.log
for args in "--rows 1250000 --k 10" "--rows 1250000 --k 1" "--rows 1250000 --k 32" "--rows 9472 --k 10"; do
  VS_LIB_PATH=$P/libvecsearch_b200_stamps.so timeout 200 python tools/scan_stamps.py $args >> $O/r02y_stamps.jsonl 2>> $O/r02y.err
done
cat $O/r02y_stamps.jsonl
g() { echo "## $1" >> $O/r02y_group.jsonl; shift; env "$@" timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 2000 >> $O/r02y_group.jsonl 2>> $O/r02y.err; }
for rep in 1 2; do
  g head VS_LIB_PATH=$P/libvecsearch_b200_head.so
  g product VS_X=1
done
cat $O/r02y_group.jsonl | cut -c1-420; tail -5 $O/r02y.err
