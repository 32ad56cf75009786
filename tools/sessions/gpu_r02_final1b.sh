#!/bin/bash
# round-2 FINAL 1-GPU session (after the L2-resident slice and the cursor selection on the query tail): full gpu suite,
# smoke (plain / under ncu), bench N=1 + reference arm, ncu launch list of the bench command, ncu --set full of the scan at
# the shard sizes bench.py reports traffic for, and one capture WITHOUT ncu's cache flush (what the L2-resident slice saves)
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_smoke_launches.csv python __graft_entry__.py smoke > $O/r02_smoke_ncu.log 2>&1; echo "smoke under ncu rc=$?" >> $O/r02_smoke_ncu.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref_n1.json 2> $O/r02_bench_ref_n1.err; echo "ref rc=$?" >> $O/r02_bench_ref_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?" >> $O/r02_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --blocks 1 --queries 8 --no-cpu-baseline --dedup-rows 200000 --group-queries 8 > $O/r02_bench_ncu.log 2>&1; echo "bench under ncu rc=$?" >> $O/r02_bench_ncu.log
for rows in 10000000 5000000 2500000 1250000; do
  timeout 400 ncu --set full --clock-control none -k regex:scan_topk -s 6 -c 1 -o $O/r02_prof_scan_$rows -f python tools/bench_scan.py --rows $rows --iters 1 > $O/r02_ncu_scan_$rows.log 2>&1
done
for rows in 1250000 10000000; do
  timeout 400 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:scan_topk -s 40 -c 4 --csv --log-file $O/r02_scan_warm_l2_$rows.csv python tools/bench_scan.py --rows $rows --iters 2 > $O/r02_scan_warm_l2_$rows.log 2>&1
done
timeout 200 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0 --queries 2000 > $O/r02_final_group_1gpu.jsonl 2>> $O/r02_final.err
timeout 200 python tools/bench_scan.py --rows 1250000,2500000,5000000,10000000 --queries 32 --iters 8 > $O/r02_final_scan.jsonl 2>> $O/r02_final.err
tail -3 $O/r02_pytest_gpu.log; tail -1 $O/r02_smoke.log; tail -1 $O/r02_smoke_ncu.log; tail -1 $O/r02_bench_n1.err; tail -1 $O/r02_bench_ncu.log; cat $O/r02_final_group_1gpu.jsonl $O/r02_final_scan.jsonl; cat $O/r02_scan_warm_l2_1250000.csv | tail -20
