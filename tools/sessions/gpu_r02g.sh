#!/bin/bash
# round-2 GPU session G (1 GPU): K2 cluster 8 vs pairs+A-groups by shard length (sustained), tensor parity tests, bench N=1
set -u
O=gpurun_out
T=multimodal-image-similarity-search_b200/libvecsearch_b200_tuning.so
S=$O/r02g_k2_threshold.jsonl
for rows in 1250000 2500000 5000000 10000000; do
  for c in 8 2; do
    VS_LIB_PATH=$T VS_TC_CLUSTER=$c timeout 300 python tools/bench_tensor.py --rows $rows --skip filter,dedup --tag "cluster<=$c" >> $S 2>> $O/r02g.err
  done
done
timeout 300 python tools/bench_tensor.py --rows 1250000 --skip filter,dedup --tag "product (adaptive)" >> $S 2>> $O/r02g.err
timeout 300 python tools/bench_tensor.py --rows 10000000 --skip filter,dedup --tag "product (adaptive)" >> $S 2>> $O/r02g.err
timeout 600 python -m pytest tests/test_gpu_tensor_parity.py tests/test_gpu_full_size.py tests/test_gpu_group.py tests/test_gpu_exchange.py -x -q > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest.log
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02g_bench_n1.json 2> $O/r02g_bench_n1.err; echo "bench rc=$?" >> $O/r02g_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02g_bench_ref.json 2> $O/r02g_bench_ref.err; echo "ref rc=$?" >> $O/r02g_bench_ref.err
tail -3 $O/r02g_pytest.log; tail -2 $O/r02g_bench_n1.err
