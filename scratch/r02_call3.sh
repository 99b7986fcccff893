#!/bin/bash
# Round 2, GPU call 3: regenerate the reference-kernel fixture (triangle order of rotate_extrude now follows
# the reference's), full GPU suite, first bench lines on the reference's own detector (both arms).
mkdir -p gpurun_out
OUT=gpurun_out/r02_call3.log
: > $OUT
python tests/golden/make_golden_gpu.py gpurun_out >> $OUT 2>&1
cp gpurun_out/ref_kernel_histories.npz tests/golden/ref_kernel_histories.npz
timeout 1500 python -m pytest tests -q -x 2>&1 | tail -30 >> $OUT
echo "=== bench reference arm" >> $OUT
timeout 1200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -c 3000 gpurun_out/r02_bench_reference.err >> $OUT; cat gpurun_out/r02_bench_reference.json >> $OUT
echo "=== bench ours" >> $OUT
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err
tail -c 3000 gpurun_out/r02_bench_ours.err >> $OUT; cat gpurun_out/r02_bench_ours.json >> $OUT
echo "=== trace" >> $OUT
timeout 600 python scratch/sweep.py "TRACE=1" "STATS=1" >> $OUT 2>&1
cat $OUT | cut -c1-2500
