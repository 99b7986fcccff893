#!/bin/bash
# Round 2, GPU call 6 (1 GPU): full GPU suite, tail-mode A/B, evidence captures (launch list, ncu --set full).
tag=r02
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1
tail -15 $out/${tag}_pytest_gpu.log
echo "=== tail A/B"
timeout 900 python scratch/sweep.py "TAIL_MODE=warp" "TAIL_MODE=warp,TRACE=1" "TAIL_MODE=lanes" "TAIL_MODE=lanes,TRACE=1" "TAIL_MODE=lanes,TAIL=700000,TRACE=1" 2>&1 | tee $out/${tag}_tail_ab.log
WORKLOAD=scint PHOTONS=10000000 timeout 600 python scratch/sweep.py "TAIL_MODE=warp" "TAIL_MODE=lanes" 2>&1 | tee -a $out/${tag}_tail_ab.log
B="python bench.py --steps 1 --warmup 3 --cpu-sample 40000"
$B > $out/${tag}_plain.json 2> $out/${tag}_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_list.log 2>&1
N="ncu --set full --clock-control none --import-source on -f"
$B > /dev/null 2>&1 && timeout 900 $N -k regex:step_intersect_kernel -c 1 -o $out/${tag}_step_intersect $B > $out/${tag}_ncu_full.log 2>&1
$B > /dev/null 2>&1 && timeout 900 $N -k regex:step_intersect_kernel --launch-skip 1 -c 1 -o $out/${tag}_step_intersect1 $B > $out/${tag}_ncu_int1.log 2>&1
$B > /dev/null 2>&1 && timeout 900 $N -k regex:propagate_tail --launch-skip 1 -c 1 -o $out/${tag}_tail $B > $out/${tag}_ncu_tail.log 2>&1
$B > /dev/null 2>&1 && timeout 900 $N -k regex:step_physics_kernel -c 1 -o $out/${tag}_physics $B > $out/${tag}_ncu_phys.log 2>&1
ls -la $out | tail -20
