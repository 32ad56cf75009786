#!/bin/bash
# round-2 GPU session C (1 GPU): full gpu test-suite, tensor-path cluster sweep (tuning build), ncu captures
set -u
O=gpurun_out
T=multimodal-image-similarity-search_b200/libvecsearch_b200_tuning.so
timeout 1100 python -m pytest tests -m gpu -x -q -s > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
for rows in 10000000 1250000; do
  for c in 8 4 2; do
    VS_LIB_PATH=$T VS_TC_CLUSTER=$c timeout 200 python tools/bench_tensor.py --rows $rows --skip filter,dedup --tag "K2 cluster<=$c" >> $O/r02c_tensor_sweep.jsonl 2>> $O/r02c_tensor_sweep.err
  done
done
VS_LIB_PATH=$T timeout 200 python tools/bench_tensor.py --rows 10000000 --skip topk,dedup --tag "K3 default" >> $O/r02c_tensor_sweep.jsonl 2>> $O/r02c_tensor_sweep.err
VS_LIB_PATH=$T timeout 200 python tools/bench_tensor.py --rows 10000000 --filters 1024 --skip topk,dedup --tag "K3 F=1024 cluster<=8" >> $O/r02c_tensor_sweep.jsonl 2>> $O/r02c_tensor_sweep.err
VS_LIB_PATH=$T VS_TC_CLUSTER=2 timeout 200 python tools/bench_tensor.py --rows 10000000 --filters 1024 --skip topk,dedup --tag "K3 F=1024 cluster<=2 (4 A groups share a slice)" >> $O/r02c_tensor_sweep.jsonl 2>> $O/r02c_tensor_sweep.err
timeout 200 python tools/bench_tensor.py --rows 2500000 --k 100 --batch 64 --skip filter,dedup --tag "K2 k=100 B=64 (4 rounds)" >> $O/r02c_tensor_sweep.jsonl 2>> $O/r02c_tensor_sweep.err
timeout 300 python tools/bench_tensor.py --rows 1250000 --skip topk,filter --dedup-rows 400000 --tag "K4" >> $O/r02c_tensor_sweep.jsonl 2>> $O/r02c_tensor_sweep.err
# ncu: K3 at 10M, K2 at 1.25M (the N=8 shard), scan at 1.25M bf16 and 1M f32 (traffic of the shapes bench.py reports)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 3 -c 1 -o $O/r02c_prof_k3 -f python tools/bench_tensor.py --rows 10000000 --skip topk,dedup --iters 1 > $O/r02c_ncu_k3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 3 -c 1 -o $O/r02c_prof_k2 -f python tools/bench_tensor.py --rows 1250000 --skip filter,dedup --iters 1 > $O/r02c_ncu_k2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:scan_topk -s 6 -c 1 -o $O/r02c_prof_scan_1250k -f python tools/bench_scan.py --rows 1250000 --iters 1 > $O/r02c_ncu_scan.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:scan_topk -s 6 -c 1 -o $O/r02c_prof_scan_f32 -f python tools/bench_scan.py --rows 1000000 --dtype f32 --iters 1 > $O/r02c_ncu_scan_f32.log 2>&1
ls -la $O/*.ncu-rep
tail -3 $O/r02c_pytest.log
