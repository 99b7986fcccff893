#!/bin/bash
# Round 2, GPU call 1: traversal-tree A/B (default / single / single + 24 stack entries), L2 access-policy
# window sizes, tail-kernel CTA shapes.  One B200.  Output: gpurun_out/r02_ab1.log
mkdir -p gpurun_out
OUT=gpurun_out/r02_ab1.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $OUT
python - >> $OUT 2>&1 <<'PY'
import ctypes, os
print('host cores', os.cpu_count())
PY
echo "=== default lib: trees x L2 window" >> $OUT
timeout 900 python scratch/sweep.py "TREE=" "L2_WINDOW_MB=16" "L2_WINDOW_MB=48" "L2_WINDOW_MB=96" "TRACE=1" \
    "TREE=single" "TREE=single,L2_WINDOW_MB=48" "TREE=single,TRACE=1" "TREE=single,STATS=1" "TREE=,STATS=1" >> $OUT 2>&1
for v in pstack24 tail768 tail512; do
  echo "=== lib_$v" >> $OUT
  if [ $v = pstack24 ]; then specs='TREE=single TREE=single,TRACE=1'; else specs='TREE= TRACE=1'; fi
  CHROMA_B200_LIB=$PWD/scratch/lib_$v.so timeout 600 python scratch/sweep.py $specs >> $OUT 2>&1
done
echo "=== pytest -m gpu" >> $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $OUT
cat $OUT
