#!/bin/bash
# Round 2, GPU call 24 (1 GPU): the record run -- GPU suite, both bench arms, the other workloads, launch list and
# ncu --set full of the hot kernels (tag r02g).
mkdir -p gpurun_out
out=gpurun_out; tag=r02g
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err
python -c "
import json
j=json.load(open('$out/${tag}_bench_ours.json')); r=json.load(open('$out/${tag}_bench_reference.json'))
print('ours value %.1f e2e %.1f | reference value %.1f e2e %.1f' % (j['value']/1e6, j['e2e']['value']/1e6, r['value']/1e6, r['e2e']['value']/1e6), j['roofline']['frac'])"
for w in rays scint pdf; do
  timeout 600 python bench.py --workload $w > $out/${tag}_bench_${w}_ours.json 2> $out/${tag}_bench_${w}_ours.err
  timeout 600 python bench.py --workload $w --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_${w}_reference.json 2> $out/${tag}_bench_${w}_reference.err
  python -c "
import json
a=json.load(open('$out/${tag}_bench_${w}_ours.json')); b=json.load(open('$out/${tag}_bench_${w}_reference.json'))
print('$w ours %.4g reference %.4g %s ratio %.2f' % (a['value'], b['value'], a['unit'], a['value']/b['value']))"
done
timeout 900 python bench.py --workload pmt29k_heavy --steps 5 --warmup 3 > $out/${tag}_bench_heavy_ours.json 2> $out/${tag}_bench_heavy_ours.err
python -c "
import json
a=json.load(open('$out/${tag}_bench_heavy_ours.json')); print('heavy value %.1f e2e %.1f M/s' % (a['value']/1e6, a['e2e']['value']/1e6), a['config']['triangles'])"
B="python bench.py --steps 1 --warmup 3 --cpu-sample 40000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_list.log 2>&1
N="ncu --set full --clock-control none --import-source on -f"
timeout 900 $N -k regex:step_intersect_kernel -c 1 -o $out/${tag}_step_intersect $B > $out/${tag}_ncu_full.log 2>&1
timeout 900 $N -k regex:step_intersect_kernel --launch-skip 1 -c 1 -o $out/${tag}_step_intersect1 $B > $out/${tag}_ncu_int1.log 2>&1
timeout 900 $N -k regex:propagate_tail --launch-skip 1 -c 1 -o $out/${tag}_tail $B > $out/${tag}_ncu_tail.log 2>&1
timeout 900 $N -k regex:step_physics_kernel -c 1 -o $out/${tag}_physics $B > $out/${tag}_ncu_phys.log 2>&1
ls -la $out/${tag}*.ncu-rep
