#!/bin/bash
# Round 2, GPU call 21 (1 GPU): cheap expansion (process4_roomy) A/B and stack size, exactness tests on the default build.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call21.log
: > $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "intersect or full_size or tree or scheduler or fixtures" 2>&1 | tail -3 >> $OUT
for v in default roomy0 p20 p28 p32 default roomy0; do
  echo "=== lib $v" >> $OUT
  if [ $v = default ]; then unset CHROMA_B200_LIB; else export CHROMA_B200_LIB=$PWD/scratch/lib_$v.so; fi
  timeout 600 python scratch/sweep.py "TRACE=1" "" 2>&1 | grep -E "SPEC|trace" >> $OUT
done
cat $OUT
