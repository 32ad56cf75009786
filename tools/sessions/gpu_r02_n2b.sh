#!/bin/bash
# 2-GPU check of bench.py after moving the one-request-at-a-time measurement in front of the NCCL rendezvous
set -u
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench rc=$?" >> $O/r02_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r02_bench_ref_n2.json 2> $O/r02_bench_ref_n2.err; echo "ref rc=$?" >> $O/r02_bench_ref_n2.err
tail -3 $O/r02_bench_n2.err; tail -2 $O/r02_bench_ref_n2.err; ls /tmp | grep vs_bench
