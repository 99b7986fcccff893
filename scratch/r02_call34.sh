#!/bin/bash
# Round 2, GPU call 34 (1 GPU): config 2 (rays) with the traversal variants of this round, same box, alternating.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call34.log
: > $OUT
for rep in 1 2; do
for v in default roomy0 tridouble old; do
  if [ $v = default ]; then unset CHROMA_B200_LIB; else export CHROMA_B200_LIB=$PWD/scratch/lib_$v.so; fi
  timeout 300 python bench.py --workload rays --steps 20 > gpurun_out/rays_$v.json 2>/dev/null
  python -c "
import json; a=json.load(open('gpurun_out/rays_$v.json')); print('$v rays/s %.4g ms %.3f e2e %.4g' % (a['value'], a['ms_per_step'], a['e2e']['value']))" >> $OUT
done
done
cat $OUT
