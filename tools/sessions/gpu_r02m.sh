#!/bin/bash
# session M (1 GPU): suspend-time hint on the long mbarrier waits of the tcgen05 kernel, same-box A/B (sustained + burst)
set -u
O=gpurun_out
D=multimodal-image-similarity-search_b200
S=$O/r02m_wait_hint.jsonl
for rep in 1 2; do
for lib in tuning hint2000 hint20000; do
  VS_LIB_PATH=$D/libvecsearch_b200_$lib.so timeout 300 python tools/bench_tensor.py --rows 10000000 --skip dedup --tag "$lib rep$rep" >> $S 2>> $O/r02m.err
  VS_LIB_PATH=$D/libvecsearch_b200_$lib.so timeout 300 python tools/bench_tensor.py --rows 1250000 --skip dedup --tag "$lib rep$rep" >> $S 2>> $O/r02m.err
done
done
for lib in tuning hint2000; do
  VS_LIB_PATH=$D/libvecsearch_b200_$lib.so timeout 300 python tools/bench_tensor.py --skip topk,filter --dedup-rows 400000 --tag "$lib" >> $S 2>> $O/r02m.err
done
