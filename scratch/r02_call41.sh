#!/bin/bash
# Round 2, GPU call 41 (1 GPU): device-block reservation for the pipeline, rays bench without allocations in its timed span.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bank_daq.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2 | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-sample 100000 > gpurun_out/r02l_bench_ours.json 2> gpurun_out/r02l_bench_ours.err || tail -20 gpurun_out/r02l_bench_ours.err
python -c "
import json
j=json.load(open('gpurun_out/r02l_bench_ours.json')); r=j['extra']['setup']['per_rank'][0]
print('value %.1f e2e %.1f M/s first yield %s gap median %.2f max %.2f' % (j['value']/1e6, j['e2e']['value']/1e6, r['first_yield_ms'], r['yield_gap_ms_median'], r['yield_gap_ms_max']), j['strong_scaling']['checksum'])"
timeout 200 python bench.py --workload rays --steps 20 > gpurun_out/r02l_bench_rays_ours.json 2>/dev/null; python -c "
import json; a=json.load(open('gpurun_out/r02l_bench_rays_ours.json')); print('rays %.4g ms %.3f max step %.2f' % (a['value'], a['ms_per_step'], max(a['extra']['ms_steps'])))"
