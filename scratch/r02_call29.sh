#!/bin/bash
# Round 2, GPU call 29 (1 GPU): GPU suite and bench line after the zero-skip upload.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02h_pytest_gpu.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_ours.json 2> gpurun_out/r02h_bench_ours.err || tail -20 gpurun_out/r02h_bench_ours.err
python -c "
import json
j=json.load(open('gpurun_out/r02h_bench_ours.json')); r=j['extra']['setup']['per_rank'][0]
print('value %.1f e2e %.1f M/s ms/step %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step']), j['e2e'], 'first yield', r['first_yield_ms'])
print(r['device_ms_per_event'])
print(r['stage_log_head'])"
python -c "import __graft_entry__ as g; g.smoke()"
