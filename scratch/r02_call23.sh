#!/bin/bash
# Round 2, GPU call 23 (1 GPU): bench line with the cheap expansion; the pipeline's stage log of the end-to-end region.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_c23.json 2> gpurun_out/r02_bench_n1_c23.err || tail -20 gpurun_out/r02_bench_n1_c23.err
python - <<P
import json
j=json.load(open('gpurun_out/r02_bench_n1_c23.json'))
r=j['extra']['setup']['per_rank'][0]
print('value %.1f e2e %.1f M/s ms/step %.3f int0 %.3f frac %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['roofline']['ms_per_launch'], j['roofline']['frac']))
print('first yield', r['first_yield_ms'], 'e2e_s', r['e2e_s'], 'loop', r['loop_s'])
for x in r['stage_log_head']: print('  ', x)
print('   ...')
for x in r['stage_log_tail']: print('  ', x)
print(j['strong_scaling'])
P
