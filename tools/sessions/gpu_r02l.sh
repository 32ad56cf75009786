#!/bin/bash
# session L (1 GPU): driver-like checks -- full gpu suite, smoke plain and UNDER ncu (kernel serialisation), bench N=1 + reference arm,
# ncu launch list of the bench command
set -u
O=gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02l_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/r02l_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02l_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02l_smoke_launches.csv python __graft_entry__.py smoke > $O/r02l_smoke_ncu.log 2>&1; echo "smoke under ncu rc=$?" >> $O/r02l_smoke_ncu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02l_bench_n1.json 2> $O/r02l_bench_n1.err; echo "bench rc=$?" >> $O/r02l_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02l_launches_bench.csv python bench.py --steps 2 --warmup 3 --blocks 1 --queries 4 --no-cpu-baseline --dedup-rows 200000 --group-queries 8 > $O/r02l_bench_ncu.log 2>&1; echo "bench under ncu rc=$?" >> $O/r02l_bench_ncu.log
tail -3 $O/r02l_pytest.log; tail -2 $O/r02l_smoke.log; tail -2 $O/r02l_smoke_ncu.log; tail -2 $O/r02l_bench_n1.err; tail -2 $O/r02l_bench_ncu.log
