#!/bin/bash
# A/B library variants through scratch/sweep.py: scratch/abs.sh "<sweep specs>" default lib_x.so ...
specs=$1; shift
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset CHROMA_B200_LIB; else export CHROMA_B200_LIB=/root/repo/scratch/$lib; fi
  echo "=== LIB $lib"
  timeout 300 python scratch/sweep.py $specs 2>&1 | grep -E "SPEC|trace|all"
done
