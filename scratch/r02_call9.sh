#!/bin/bash
# Round 2, GPU call 9 (1 GPU): kernel variant A/B (no-allocate triangle loads, lane prefetch, tail prefetch off,
# 16-entry stack), ray sorting, then the GPU suite and a bench line.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call9.log
: > $OUT
echo "=== default lib" >> $OUT
timeout 900 python scratch/sweep.py "TRACE=1" "SORT=1000000,TRACE=1" "SORT=100000,TRACE=1" "TAIL_MODE=warp,TRACE=1" "L2_WINDOW_MB=0" >> $OUT 2>&1
for v in trina lanepf nopf pstack16; do
  echo "=== lib_$v" >> $OUT
  CHROMA_B200_LIB=$PWD/scratch/lib_$v.so timeout 600 python scratch/sweep.py "TRACE=1" "" >> $OUT 2>&1
done
grep -E "SPEC|trace|===|Error" $OUT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; python -c "
import json; j=json.load(open('gpurun_out/r02_bench_n1_b.json')); print('value %.1f e2e %.1f' % (j['value']/1e6, j['e2e']['value']/1e6), j['strong_scaling'])"
