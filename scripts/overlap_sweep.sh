python scripts/overlap_probe.py 16 1 | tail -1
for nb in 8 6 5 4; do echo "== trace blocks/SM $nb"; PBRGPU_TRACE_BLOCKS_PER_SM=$nb python scripts/overlap_probe.py 8 2 | tail -1; PBRGPU_TRACE_BLOCKS_PER_SM=$nb python scripts/overlap_probe.py 16 2 | tail -1; done
PBRGPU_TRACE_BLOCKS_PER_SM=4 python scripts/overlap_probe.py 16 3 | tail -1
