#!/bin/bash
# round-2 FINAL 1-GPU session, part 2 (part 1 = tools/sessions/gpu_r02_final1b.sh: 113 gpu tests passed, smoke ok plain and under ncu;
# its outputs exceeded the 64 MiB return limit): bench N=1, launch list of the bench command, one --set full capture of the
# scan, and single-metric captures WITHOUT ncu's cache flush / replay (dram bytes of back-to-back scans)
set -u
O=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?" >> $O/r02_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --blocks 1 --queries 8 --no-cpu-baseline --dedup-rows 200000 --group-queries 8 > $O/r02_bench_ncu.log 2>&1; echo "bench under ncu rc=$?" >> $O/r02_bench_ncu.log
timeout 400 ncu --set full --clock-control none -k regex:scan_topk -s 6 -c 1 -o $O/r02_prof_scan_10000000 -f python tools/bench_scan.py --rows 10000000 --iters 1 > $O/r02_ncu_scan_10000000.log 2>&1
for rows in 1250000 10000000; do
  for m in dram__bytes_read.sum lts__t_sector_hit_rate.pct; do
    timeout 400 ncu --cache-control none --clock-control none --metrics $m -k regex:scan_topk -s 40 -c 6 --csv --log-file $O/r02_scan_warm_l2_${rows}_$m.csv python tools/bench_scan.py --rows $rows --iters 2 > $O/r02_scan_warm_l2.log 2>&1
  done
done
timeout 300 python -m pytest tests/test_gpu_scan_parity.py -x -q -m gpu > $O/r02_pytest_scan.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_scan.log
tail -1 $O/r02_bench_n1.err; tail -1 $O/r02_bench_ncu.log; tail -2 $O/r02_pytest_scan.log
grep -h "scan_topk" $O/r02_scan_warm_l2_*.csv | awk -F'","' '{print $13, $15}' | sort | uniq -c | head -30
