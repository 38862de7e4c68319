#!/bin/bash
for th in 1 6 12; do
  PBRGPU_TRI_THRESHOLD=$th PBRGPU_WAVE_LOG2=23 python scripts/quick_perf.py 467 4 > gpurun_out/plain_th$th.log 2>&1 && \
  PBRGPU_TRI_THRESHOLD=$th PBRGPU_WAVE_LOG2=23 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active --clock-control none -k regex:k_trace -s 1 -c 2 --csv --log-file gpurun_out/th$th.csv python scripts/quick_perf.py 467 4 > /dev/null 2>&1
  grep "^ms" gpurun_out/plain_th$th.log | sed -n 2p
done
