#!/bin/bash
# Round 2, GPU call 2: parity suite on the rewritten physics, then timing of the same A/B set.
mkdir -p gpurun_out
OUT=gpurun_out/r02_ab2.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 >> $OUT
echo "=== default lib" >> $OUT
timeout 900 python scratch/sweep.py "TREE=" "TRACE=1" "L2_WINDOW_MB=16" "TREE=single,L2_WINDOW_MB=16" >> $OUT 2>&1
for v in pstack24 p24t768 tail768; do
  echo "=== lib_$v" >> $OUT
  if [ $v = tail768 ]; then specs='TREE= TRACE=1'; else specs='TREE=single TREE=single,L2_WINDOW_MB=16 TREE=single,L2_WINDOW_MB=16,TRACE=1'; fi
  CHROMA_B200_LIB=$PWD/scratch/lib_$v.so timeout 600 python scratch/sweep.py $specs >> $OUT 2>&1
done
echo "=== scint" >> $OUT
WORKLOAD=scint PHOTONS=10000000 timeout 600 python scratch/sweep.py "TREE=" "TRACE=1" >> $OUT 2>&1
cat $OUT
