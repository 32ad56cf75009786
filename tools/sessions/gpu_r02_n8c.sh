#!/bin/bash
# round-2 final 8-GPU session after the scan-tail work: functional worker, bench N=8 and N=4
set -u
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tests/p2p_worker.py > $O/r02_p2p_worker_n8.log 2>&1; echo "p2p rc=$?" >> $O/r02_p2p_worker_n8.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench rc=$?" >> $O/r02_bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 4 --steps 20 --warmup 5 --dedup-rows 0 > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; echo "bench rc=$?" >> $O/r02_bench_n4.err
timeout 120 python tools/bench_group.py --rows-per-gpu 1250000 --queries 2000 > $O/r02_final_group_8gpu.jsonl 2>> $O/r02_bench_n8.err
tail -2 $O/r02_p2p_worker_n8.log; tail -2 $O/r02_bench_n8.err; tail -2 $O/r02_bench_n4.err; cat $O/r02_final_group_8gpu.jsonl | cut -c1-420
