#!/bin/bash
# round-2 multi-GPU session (8 GPUs): functional worker (fused exchange == NCCL == oracle, ShardedIndex service), bench N=8 and N=4
set -u
O=gpurun_out
nvidia-smi topo -m > $O/r02_n8_topo.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tests/p2p_worker.py > $O/r02_p2p_worker_n8.log 2>&1; echo "p2p rc=$?" >> $O/r02_p2p_worker_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench rc=$?" >> $O/r02_bench_n8.err
timeout 300 python -m pytest tests/test_gpu_group.py -x -q > $O/r02_n8_group_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02_n8_group_tests.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 4 --steps 20 --warmup 5 --dedup-rows 0 > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; echo "bench rc=$?" >> $O/r02_bench_n4.err
tail -2 $O/r02_p2p_worker_n8.log; tail -2 $O/r02_bench_n8.err; tail -3 $O/r02_n8_group_tests.log; tail -2 $O/r02_bench_n4.err
