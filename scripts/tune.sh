#!/bin/bash
# development aid: sweep traversal knobs on the bench workload
for leaf in 3 2 1; do for th in 1 8 12 16 20 24; do
  echo "== leaf $leaf thresh $th"; PBRGPU_MAX_LEAF=$leaf PBRGPU_TRI_THRESHOLD=$th python scripts/quick_perf.py 467 4 2>&1 | grep "^ms" | tail -2
done; done
