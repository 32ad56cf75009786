#!/bin/bash
# round-2 FINAL 1-GPU session: full gpu suite, smoke (plain / under ncu / under compute-sanitizer), bench N=1 + reference arm,
# ncu launch list of the bench command, ncu --set full of the scan at the shard sizes bench.py reports traffic for
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_smoke_launches.csv python __graft_entry__.py smoke > $O/r02_smoke_ncu.log 2>&1; echo "smoke under ncu rc=$?" >> $O/r02_smoke_ncu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/r02_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?" >> $O/r02_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q "tests/test_gpu_group.py::test_remove_rows_is_one_compaction_and_keeps_answers" "tests/test_gpu_group.py::test_raw_slab_roundtrip_is_bit_exact" "tests/test_gpu_tensor_parity.py::test_tensor_topk_rounds_and_a_groups[90-512-16-64]" "tests/test_gpu_tensor_parity.py::test_tensor_rounds_with_pre_filter" "tests/test_gpu_collection.py" > $O/r02_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?" >> $O/r02_memcheck_tests.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref_n1.json 2> $O/r02_bench_ref_n1.err; echo "ref rc=$?" >> $O/r02_bench_ref_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?" >> $O/r02_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --blocks 1 --queries 8 --no-cpu-baseline --dedup-rows 200000 --group-queries 8 > $O/r02_bench_ncu.log 2>&1; echo "bench under ncu rc=$?" >> $O/r02_bench_ncu.log
for rows in 10000000 5000000 2500000; do
  timeout 400 ncu --set full --clock-control none -k regex:scan_topk -s 6 -c 1 -o $O/r02_prof_scan_$rows -f python tools/bench_scan.py --rows $rows --iters 1 > $O/r02_ncu_scan_$rows.log 2>&1
done
tail -3 $O/r02_pytest_gpu.log; tail -1 $O/r02_smoke.log; tail -1 $O/r02_smoke_ncu.log; tail -2 $O/r02_memcheck_smoke.log; tail -3 $O/r02_memcheck_tests.log; tail -1 $O/r02_bench_n1.err; tail -1 $O/r02_bench_ncu.log
