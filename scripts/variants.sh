#!/bin/bash
echo "== base"; python scripts/quick_perf.py 467 16 2>&1 | grep "^ms" | tail -2
for v in "$@"; do
  echo "== $v"; PBRGPU_LIB=$PWD/build_variants/$v.so python scripts/quick_perf.py 467 16 2>&1 | grep "^ms" | tail -2
done
