# Round-2 profiling pass (run under gpurun on one GPU; every ncu command follows a plain run of the same program that exited 0).
set -x
M=$(python scripts/ncu_counters.py)
python scripts/one_wave.py > gpurun_out/one_wave.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/ncu_counters_trace.csv python scripts/one_wave.py > gpurun_out/ncu1.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_shade --csv --log-file gpurun_out/ncu_counters_shade.csv python scripts/one_wave.py > gpurun_out/ncu2.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/r02_k_trace_full -f python scripts/one_wave.py > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 1 -o gpurun_out/r02_k_shade_full -f python scripts/one_wave.py > gpurun_out/ncu5.log 2>&1
PBRGPU_LANE_LOG=1 python scripts/quick_perf.py 467 16 > gpurun_out/lane_log.txt 2>&1
ls -la gpurun_out
