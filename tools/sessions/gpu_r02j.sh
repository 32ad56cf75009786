#!/bin/bash
# session J (1 GPU): parity after the selection-based tail merge; per-request latency again; pipelined scan rate unchanged?
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan_parity.py tests/test_gpu_exchange.py tests/test_gpu_group.py tests/test_gpu_collection.py tests/test_gpu_full_size.py tests/test_property.py -x -q > $O/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02j_pytest.log
for rows in 32 9472 1250000 10000000; do
  timeout 200 python tools/bench_group.py --rows-per-gpu $rows --devices 0 --queries 1500 >> $O/r02j_group.jsonl 2>> $O/r02j.err
done
timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 1500 --k 32 >> $O/r02j_group.jsonl 2>> $O/r02j.err
timeout 200 python tools/bench_scan.py --rows 1250000,10000000 >> $O/r02j_scan.jsonl 2>> $O/r02j.err
tail -3 $O/r02j_pytest.log; cat $O/r02j_group.jsonl $O/r02j_scan.jsonl
