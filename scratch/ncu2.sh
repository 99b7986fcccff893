#!/bin/bash
# second-step intersect launch and the real (second) tail launch under ncu --set full; per-step stats
out=gpurun_out; tag=${1:-r01s3}
python scratch/sweep.py "STATS=1,TRACE=1" > $out/${tag}_stats.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_intersect_kernel --launch-skip 1 -c 1 \
   -o $out/${tag}_step_intersect1 -f python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_int1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_tail_kernel --launch-skip 1 -c 1 \
   -o $out/${tag}_tail -f python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_tail.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_physics_kernel -c 1 \
   -o $out/${tag}_physics -f python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_phys.log 2>&1
tail -30 $out/${tag}_stats.log
