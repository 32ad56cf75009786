#!/bin/bash
# round-2 GPU session F (1 GPU): K2 cluster size under SUSTAINED load (>= 0.3 s timed loops), new epilogue, same box as the previous-commit library
set -u
O=gpurun_out
T=multimodal-image-similarity-search_b200/libvecsearch_b200_tuning.so
P=multimodal-image-similarity-search_b200/libvecsearch_b200_prev.so
S=$O/r02f_k2_sustained.jsonl
for rows in 10000000 1250000; do
  VS_LIB_PATH=$P timeout 300 python tools/bench_tensor.py --rows $rows --skip filter,dedup --tag "prev (cluster 8, old epilogue)" >> $S 2>> $O/r02f.err
  for c in 8 4 2; do
    VS_LIB_PATH=$T VS_TC_CLUSTER=$c timeout 300 python tools/bench_tensor.py --rows $rows --skip filter,dedup --tag "new epilogue, cluster<=$c" >> $S 2>> $O/r02f.err
  done
done
VS_LIB_PATH=$T VS_TC_CLUSTER=8 timeout 300 python tools/bench_tensor.py --rows 2500000 --dim 768 --skip filter,dedup --tag "dim 768 new epilogue, cluster<=8" >> $S 2>> $O/r02f.err
VS_LIB_PATH=$T VS_TC_CLUSTER=2 timeout 300 python tools/bench_tensor.py --rows 2500000 --dim 768 --skip filter,dedup --tag "dim 768 new epilogue, cluster<=2" >> $S 2>> $O/r02f.err
VS_LIB_PATH=$T VS_TC_CLUSTER=8 timeout 300 python tools/bench_tensor.py --rows 10000000 --batch 256 --skip filter,dedup --tag "B=256 cluster<=8 (=2)" >> $S 2>> $O/r02f.err
VS_LIB_PATH=$T VS_TC_CLUSTER=8 timeout 300 python tools/bench_tensor.py --rows 10000000 --batch 2048 --skip filter,dedup --tag "B=2048 cluster<=8, 2 A groups" >> $S 2>> $O/r02f.err
