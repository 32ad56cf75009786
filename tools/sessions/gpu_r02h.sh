#!/bin/bash
# round-2 GPU session H (1 GPU): scan parity after the prefetching tail merge, per-request latency of the group path by shard size
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_scan_parity.py tests/test_gpu_exchange.py tests/test_gpu_group.py tests/test_gpu_collection.py tests/test_c_host.py -x -q > $O/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02h_pytest.log
for rows in 1250000 2500000 10000000; do
  timeout 200 python tools/bench_group.py --rows-per-gpu $rows --devices 0 >> $O/r02h_group.jsonl 2>> $O/r02h.err
done
timeout 200 python tools/bench_scan.py --rows 1250000,10000000 >> $O/r02h_scan.jsonl 2>> $O/r02h.err
timeout 200 python tools/bench_scan.py --rows 1000000 --dtype f32 >> $O/r02h_scan.jsonl 2>> $O/r02h.err
tail -3 $O/r02h_pytest.log; cat $O/r02h_group.jsonl; cat $O/r02h_scan.jsonl
