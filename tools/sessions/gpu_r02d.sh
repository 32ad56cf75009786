#!/bin/bash
# round-2 GPU session D (1 GPU): full gpu test-suite, K3 after the epilogue rewrite, K4 cluster sweep, ncu of K3
set -u
O=gpurun_out
T=multimodal-image-similarity-search_b200/libvecsearch_b200_tuning.so
timeout 1100 python -m pytest tests -m gpu -x -q -s > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log
S=$O/r02d_tensor_sweep.jsonl
timeout 200 python tools/bench_tensor.py --rows 10000000 --skip dedup --tag "product lib: K2 (pairs + A groups) / K3 (packed epilogue)" >> $S 2>> $O/r02d_tensor.err
timeout 200 python tools/bench_tensor.py --rows 1250000 --skip dedup --tag "product lib, N=8 shard" >> $S 2>> $O/r02d_tensor.err
timeout 200 python tools/bench_tensor.py --rows 10000000 --filters 1024 --skip topk,dedup --tag "K3 F=1024" >> $S 2>> $O/r02d_tensor.err
for c in 8 4 2; do
  VS_LIB_PATH=$T VS_TC_CLUSTER=$c timeout 300 python tools/bench_tensor.py --skip topk,filter --dedup-rows 400000 --tag "K4 cluster<=$c" >> $S 2>> $O/r02d_tensor.err
done
VS_LIB_PATH=$T VS_TC_CLUSTER=2 timeout 300 python tools/bench_tensor.py --skip topk,filter --dedup-rows 400000 --dedup-dim 512 --tag "K4 dim 512 cluster<=2" >> $S 2>> $O/r02d_tensor.err
VS_LIB_PATH=$T VS_TC_CLUSTER=8 timeout 300 python tools/bench_tensor.py --skip topk,filter --dedup-rows 400000 --dedup-dim 512 --tag "K4 dim 512 cluster<=8" >> $S 2>> $O/r02d_tensor.err
timeout 200 python tools/bench_tensor.py --rows 2500000 --k 100 --batch 64 --skip filter,dedup --tag "K2 k=100 B=64 (4 rounds)" >> $S 2>> $O/r02d_tensor.err
timeout 200 python tools/bench_tensor.py --rows 2500000 --dim 768 --skip dedup --tag "dim 768 (streamed A)" >> $S 2>> $O/r02d_tensor.err
timeout 200 python tools/bench_scan.py --rows 10000000,1250000 >> $O/r02d_scan.jsonl 2>> $O/r02d_tensor.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 3 -c 1 -o $O/r02d_prof_k3 -f python tools/bench_tensor.py --rows 10000000 --skip topk,dedup --iters 1 > $O/r02d_ncu_k3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 3 -c 1 -o $O/r02d_prof_k2 -f python tools/bench_tensor.py --rows 1250000 --skip filter,dedup --iters 1 > $O/r02d_ncu_k2.log 2>&1
tail -3 $O/r02d_pytest.log
