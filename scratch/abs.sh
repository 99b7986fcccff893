#!/bin/bash
# A/B library variants through scratch/sweep.py (kernel-only timing + one traced run): scratch/abs.sh default lib_x.so ...
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset CHROMA_B200_LIB; else export CHROMA_B200_LIB=/root/repo/scratch/$lib; fi
  echo "=== LIB $lib"
  timeout 300 python scratch/sweep.py "TRACE=1" 2>&1 | grep -E "SPEC|trace|all"
done
