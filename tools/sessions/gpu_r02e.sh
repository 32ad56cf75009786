#!/bin/bash
# round-2 GPU session E (1 GPU): same-box A/B of the tensor kernels (previous commit's library vs this one), full gpu suite
set -u
O=gpurun_out
P=multimodal-image-similarity-search_b200/libvecsearch_b200_prev.so
S=$O/r02e_tensor_ab.jsonl
for rep in 1 2; do
  for lib in prev new; do
    if [ $lib = prev ]; then export VS_LIB_PATH=$P; else unset VS_LIB_PATH; fi
    timeout 300 python tools/bench_tensor.py --rows 10000000 --skip dedup --tag "$lib rep$rep" >> $S 2>> $O/r02e_tensor.err
    timeout 300 python tools/bench_tensor.py --rows 1250000 --skip dedup --tag "$lib rep$rep N=8 shard" >> $S 2>> $O/r02e_tensor.err
    timeout 300 python tools/bench_tensor.py --skip topk,filter --dedup-rows 400000 --tag "$lib rep$rep" >> $S 2>> $O/r02e_tensor.err
  done
done
unset VS_LIB_PATH
timeout 300 python tools/bench_tensor.py --rows 10000000 --filters 1024 --skip topk,dedup --tag "new K3 F=1024" >> $S 2>> $O/r02e_tensor.err
timeout 300 python tools/bench_tensor.py --rows 2500000 --k 100 --batch 64 --skip filter,dedup --tag "new K2 k=100 B=64 (4 rounds)" >> $S 2>> $O/r02e_tensor.err
timeout 1100 python -m pytest tests -m gpu -x -q -s > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02e_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 3 -c 1 -o $O/r02e_prof_k3 -f python tools/bench_tensor.py --rows 10000000 --skip topk,dedup --iters 1 > $O/r02e_ncu_k3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 3 -c 1 -o $O/r02e_prof_k2 -f python tools/bench_tensor.py --rows 1250000 --skip filter,dedup --iters 1 > $O/r02e_ncu_k2.log 2>&1
tail -3 $O/r02e_pytest.log
