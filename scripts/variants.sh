#!/bin/bash
# usage: variants.sh <variant.so names under build_variants/ ...>  - per-kernel-group times of the bench workload
echo "== base"; python scripts/perf_timers.py 2>&1 | tail -2
for v in "$@"; do
  echo "== $v"; PBRGPU_LIB=$PWD/build_variants/$v.so python scripts/perf_timers.py 2>&1 | tail -2
done
