#!/bin/bash
# development aid: sweep traversal knobs on the bench workload
run() { echo "== $*"; env "$@" python scripts/quick_perf.py 467 8 2>&1 | grep "^ms" | sed -n 2p; }
run A=1
run PBRGPU_FETCH_MIN_IDLE=4
run PBRGPU_FETCH_MIN_IDLE=12
run PBRGPU_FETCH_MIN_IDLE=16
run PBRGPU_FETCH_MIN_IDLE=24
run PBRGPU_WAVE_LOG2=21
run PBRGPU_WAVE_LOG2=23
run PBRGPU_WAVE_LOG2=24
run PBRGPU_TRACE_BLOCKS_PER_SM=4
run PBRGPU_TRACE_BLOCKS_PER_SM=6
run PBRGPU_TRACE_BLOCKS_PER_SM=7
