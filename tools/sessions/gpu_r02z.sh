#!/bin/bash
# session Z (2 GPUs): does an idle CUDA context of ANOTHER process on the same GPUs slow the group's launches?
set -u
O=gpurun_out
timeout 120 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0,1 --queries 1500 > $O/r02z.jsonl 2>> $O/r02z.err
python - <<'PY' &
import time, torch
for d in (0, 1):
    torch.zeros(1, device=f"cuda:{d}")
torch.cuda.synchronize()
time.sleep(45)
PY
BG=$!
sleep 12
timeout 120 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0,1 --queries 1500 >> $O/r02z.jsonl 2>> $O/r02z.err
kill $BG 2>/dev/null; wait $BG 2>/dev/null
OMP_NUM_THREADS=1 timeout 120 python tools/bench_group.py --rows-per-gpu 1250000 --devices 0,1 --queries 1500 >> $O/r02z.jsonl 2>> $O/r02z.err
cat $O/r02z.jsonl | cut -c1-400; tail -3 $O/r02z.err
