#!/bin/bash
# session K (1 GPU): isolated-request latency vs CTAs per SM / ring depth (tuning build)
set -u
O=gpurun_out
T=multimodal-image-similarity-search_b200/libvecsearch_b200_tuning.so
run() { echo "## $*" >> $O/r02k_group.jsonl; env "$@" VS_LIB_PATH=$T timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 2000 >> $O/r02k_group.jsonl 2>> $O/r02k.err; }
run VS_SCAN_CTAS_PER_SM=2
run VS_SCAN_CTAS_PER_SM=1 VS_SCAN_SMEM_KB=76
run VS_SCAN_CTAS_PER_SM=1 VS_SCAN_SMEM_KB=110
run VS_SCAN_CTAS_PER_SM=1 VS_SCAN_SMEM_KB=140
run VS_SCAN_CTAS_PER_SM=1 VS_SCAN_SMEM_KB=200
run VS_SCAN_CTAS_PER_SM=2 VS_SCAN_SMEM_KB=110
run VS_SCAN_CTAS_PER_SM=3 VS_SCAN_SMEM_KB=72
cat $O/r02k_group.jsonl
