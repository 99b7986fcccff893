#!/bin/bash
# One gpurun call: GPU parity tests, both bench arms, ncu launch list, ncu --set full of the hot kernels.
# usage (from the repo root on the box): bash scratch/round_gpu.sh [tag]
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
timeout 900 python bench.py > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err
rc=$?
cat $out/${tag}_bench_ours.json | cut -c1-300
if [ $rc -eq 0 ]; then
  B="python bench.py --steps 1 --warmup 3 --cpu-sample 40000"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_list.log 2>&1
  N="ncu --set full --clock-control none --import-source on -f"
  timeout 900 $N -k regex:step_intersect_kernel -c 1 -o $out/${tag}_step_intersect $B > $out/${tag}_ncu_full.log 2>&1
  timeout 900 $N -k regex:step_intersect_kernel --launch-skip 1 -c 1 -o $out/${tag}_step_intersect1 $B > $out/${tag}_ncu_int1.log 2>&1
  timeout 900 $N -k regex:propagate_tail_kernel --launch-skip 1 -c 1 -o $out/${tag}_tail $B > $out/${tag}_ncu_tail.log 2>&1
  timeout 900 $N -k regex:step_physics_kernel -c 1 -o $out/${tag}_physics $B > $out/${tag}_ncu_phys.log 2>&1
fi
timeout 600 python bench.py --workload rays > $out/${tag}_bench_rays_ours.json 2> $out/${tag}_bench_rays_ours.err
timeout 600 python bench.py --workload rays --impl reference > $out/${tag}_bench_rays_reference.json 2> $out/${tag}_bench_rays_reference.err
cat $out/${tag}_bench_rays_ours.json | cut -c1-200
ls -la $out | tail -24
