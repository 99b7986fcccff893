#!/bin/bash
# Round 2, GPU call 20 (1 GPU): GPU suite after the evidx reordering, bench line, launch list, ncu --set full of the hot kernels.
mkdir -p gpurun_out
out=gpurun_out; tag=r02f
timeout 1500 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest_gpu.log 2>&1; tail -4 $out/${tag}_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err
python -c "
import json; j=json.load(open('$out/${tag}_bench_ours.json')); print('value %.1f e2e %.1f' % (j['value']/1e6, j['e2e']['value']/1e6), j['roofline']['frac'], j['cpu_baseline'])"
B="python bench.py --steps 1 --warmup 3 --cpu-sample 40000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_list.log 2>&1
N="ncu --set full --clock-control none --import-source on -f"
timeout 900 $N -k regex:step_intersect_kernel -c 1 -o $out/${tag}_step_intersect $B > $out/${tag}_ncu_full.log 2>&1
timeout 900 $N -k regex:step_intersect_kernel --launch-skip 1 -c 1 -o $out/${tag}_step_intersect1 $B > $out/${tag}_ncu_int1.log 2>&1
timeout 900 $N -k regex:propagate_tail --launch-skip 1 -c 1 -o $out/${tag}_tail $B > $out/${tag}_ncu_tail.log 2>&1
timeout 900 $N -k regex:step_physics_kernel -c 1 -o $out/${tag}_physics $B > $out/${tag}_ncu_phys.log 2>&1
ls -la $out/*.ncu-rep
