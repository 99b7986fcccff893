#!/bin/bash
# Round 2, GPU call 7: prefetch A/B in the tail (both tail modes) and in the lane-per-ray kernel.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call7.log
: > $OUT
echo "=== default lib (tail prefetch on)" >> $OUT
timeout 900 python scratch/sweep.py "TAIL_MODE=warp,TRACE=1" "TAIL_MODE=lanes,TRACE=1" "TAIL_MODE=lanes" "TAIL_MODE=warp" >> $OUT 2>&1
echo "=== lib_nopf" >> $OUT
CHROMA_B200_LIB=$PWD/scratch/lib_nopf.so timeout 600 python scratch/sweep.py "TAIL_MODE=warp,TRACE=1" "TAIL_MODE=lanes,TRACE=1" >> $OUT 2>&1
echo "=== lib_lanepf" >> $OUT
CHROMA_B200_LIB=$PWD/scratch/lib_lanepf.so timeout 600 python scratch/sweep.py "TAIL_MODE=lanes,TRACE=1" "TAIL_MODE=lanes" >> $OUT 2>&1
grep -E "SPEC|trace|===" $OUT
