#!/bin/bash
for v in t8 t10 s3 t8s3; do
  echo "== $v"; PBRGPU_LIB=$PWD/build_variants/$v.so PBRGPU_WAVE_LOG2=24 python scripts/quick_perf.py 467 8 2>&1 | grep "^ms" | sed -n 2p
done
echo "== t8 blocks=10 smem"; PBRGPU_LIB=$PWD/build_variants/t10.so PBRGPU_WAVE_LOG2=24 PBRGPU_TRACE_BLOCKS_PER_SM=12 python scripts/quick_perf.py 467 8 2>&1 | grep "^ms" | sed -n 2p
echo "== t8 wave25"; PBRGPU_LIB=$PWD/build_variants/t8.so PBRGPU_WAVE_LOG2=25 python scripts/quick_perf.py 467 16 2>&1 | grep "^ms" | sed -n 2p
echo "== t8 wave24 16spp"; PBRGPU_LIB=$PWD/build_variants/t8.so PBRGPU_WAVE_LOG2=24 python scripts/quick_perf.py 467 16 2>&1 | grep "^ms" | sed -n 2p
