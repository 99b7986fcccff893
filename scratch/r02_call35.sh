#!/bin/bash
# Round 2, GPU call 35 (1 GPU): config 2 (rays), both arms, with the warm-up that waits for the clocks (two processes each).
mkdir -p gpurun_out
for rep in 1 2; do
for impl in ours reference; do
  timeout 300 python bench.py --workload rays --impl $impl --steps 20 > gpurun_out/r02j_bench_rays_${impl}_$rep.json 2> gpurun_out/r02j_bench_rays_${impl}_$rep.err || tail -5 gpurun_out/r02j_bench_rays_${impl}_$rep.err
  python -c "
import json; a=json.load(open('gpurun_out/r02j_bench_rays_${impl}_$rep.json')); print('$impl rep $rep: rays/s %.4g ms %.3f e2e %.4g warm-up calls %s' % (a['value'], a['ms_per_step'], a['e2e']['value'], a.get('warmup')))"
done
done
