#!/bin/bash
# Round 2, GPU call 38 (1 GPU): HEAD -- GPU suite, smoke, default bench invocation (no flags), reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02k_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02k_pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/r02k_bench_default.json 2> gpurun_out/r02k_bench_default.err || tail -20 gpurun_out/r02k_bench_default.err
python -c "
import json
j=json.load(open('gpurun_out/r02k_bench_default.json'))
print('default run: value %.1f e2e %.1f M/s steps %d warmup %d launches %d clocks %s' % (j['value']/1e6, j['e2e']['value']/1e6, j['steps'], j['warmup'], j['gpu_launches'], j['clocks']))
print(sorted(j.keys()))"
wc -l gpurun_out/r02k_bench_default.json
