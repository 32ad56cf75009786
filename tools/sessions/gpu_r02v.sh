#!/bin/bash
# session V (1 GPU): phase marks of one isolated request (stamps build)
set -u
O=gpurun_out
L=multimodal-image-similarity-search_b200/libvecsearch_b200_stamps.so
for args in "--rows 1250000 --k 10" "--rows 1250000 --k 1" "--rows 9472 --k 10" "--rows 32 --k 10" "--rows 10000000 --k 10"; do
  VS_LIB_PATH=$L timeout 200 python tools/scan_stamps.py $args >> $O/r02v_stamps.jsonl 2>> $O/r02v.err
done
cat $O/r02v_stamps.jsonl; tail -5 $O/r02v.err
