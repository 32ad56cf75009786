#!/bin/bash
# group tests + smoke after bounding the completion-flag poll (host-side change in csrc/group.cu)
set -u
O=gpurun_out
timeout 130 python -m pytest tests/test_gpu_group.py -x -q -m gpu > $O/r02_last_group_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02_last_group_tests.log
timeout 60 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
tail -3 $O/r02_last_group_tests.log; tail -1 $O/r02_smoke.log
