#!/bin/bash
# Round 2, GPU call 25 (1 GPU): predicated-add form of the cheap expansion vs the select form (lib_pipe2 / lib_nopipe).
mkdir -p gpurun_out
OUT=gpurun_out/r02_call26.log
: > $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "intersect or full_size or tree or scheduler or fixtures" 2>&1 | tail -3 >> $OUT
for v in default leafpf1 leafpf2 default leafpf1 leafpf2; do
  echo "=== lib $v" >> $OUT
  if [ $v = default ]; then unset CHROMA_B200_LIB; else export CHROMA_B200_LIB=$PWD/scratch/lib_$v.so; fi
  timeout 600 python scratch/sweep.py "TRACE=1" "" 2>&1 | grep -E "SPEC|trace" >> $OUT
done
cat $OUT
