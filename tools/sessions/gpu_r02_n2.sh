#!/bin/bash
# round-2 2-GPU session: functional worker, group tests on distinct GPUs, bench N=2
set -u
O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/p2p_worker.py > $O/r02_p2p_worker_n2.log 2>&1; echo "p2p rc=$?" >> $O/r02_p2p_worker_n2.log
timeout 300 python -m pytest tests/test_gpu_group.py tests/test_gpu_exchange.py -x -q > $O/r02_n2_group_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02_n2_group_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench rc=$?" >> $O/r02_bench_n2.err
timeout 200 python tools/bench_scan.py --rows 1250000,5000000 >> $O/r02_n2_scan.jsonl 2>> $O/r02_bench_n2.err
tail -2 $O/r02_p2p_worker_n2.log; tail -3 $O/r02_n2_group_tests.log; tail -2 $O/r02_bench_n2.err
