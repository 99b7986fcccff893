#!/bin/bash
# Round 2, GPU call 33 (1 GPU): final state -- GPU suite, smoke, bench line (tag r02i).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02i_pytest_gpu.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench_ours.json 2> gpurun_out/r02i_bench_ours.err || tail -20 gpurun_out/r02i_bench_ours.err
python -c "
import json
j=json.load(open('gpurun_out/r02i_bench_ours.json')); r=j['extra']['setup']['per_rank'][0]
print('value %.1f e2e %.1f M/s ms/step %.3f frac %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['roofline']['frac']), j['e2e'], 'first yield', r['first_yield_ms'])
print([x['frac'] for x in j['rooflines']], j['cpu_baseline']['value'], j['strong_scaling']['seconds'], j['strong_scaling']['checksum'])"
timeout 300 python bench.py --workload rays --steps 5 > gpurun_out/r02i_bench_rays_ours.json 2>/dev/null; python -c "
import json; a=json.load(open('gpurun_out/r02i_bench_rays_ours.json')); print('rays %.4g e2e %.4g' % (a['value'], a['e2e']['value']))"
