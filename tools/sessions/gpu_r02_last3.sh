#!/bin/bash
# last GPU call of the round: bitonic merge tree on the query tail (k <= 16) -- parity tests, phase marks, same-box A/B
set -u
O=gpurun_out
P=multimodal-image-similarity-search_b200
timeout 110 python -m pytest tests/test_gpu_scan_parity.py tests/test_gpu_exchange.py tests/test_gpu_group.py -x -q -m gpu > $O/r02_last3_tests.log 2>&1; echo "tests rc=$?" >> $O/r02_last3_tests.log
VS_LIB_PATH=$P/libvecsearch_b200_stamps.so timeout 40 python tools/scan_stamps.py --rows 1250000 --k 10 --queries 200 > $O/r02_last3_stamps.jsonl 2>> $O/r02_last3.err
VS_LIB_PATH=$P/libvecsearch_b200_prev.so timeout 40 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 1500 > $O/r02_last3_group.jsonl 2>> $O/r02_last3.err
timeout 40 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 1500 >> $O/r02_last3_group.jsonl 2>> $O/r02_last3.err
tail -3 $O/r02_last3_tests.log; cat $O/r02_last3_stamps.jsonl; cut -c1-330 $O/r02_last3_group.jsonl; tail -3 $O/r02_last3.err
