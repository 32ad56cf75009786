#!/bin/bash
# last check of the round: default bench.py invocation (what the driver runs at N=1) + smoke with the committed tree
set -u
O=gpurun_out
timeout 200 python bench.py > $O/r02_bench_default_n1.json 2> $O/r02_bench_default_n1.err; echo "bench rc=$?" >> $O/r02_bench_default_n1.err
timeout 100 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
tail -2 $O/r02_bench_default_n1.err; tail -2 $O/r02_smoke.log
