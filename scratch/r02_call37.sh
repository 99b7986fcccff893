#!/bin/bash
mkdir -p gpurun_out
BENCH_DEBUG=1 timeout 300 python bench.py --workload rays --steps 10 > gpurun_out/r02_call37.json 2> gpurun_out/r02_call37.err
grep -E "warm-up|timed" gpurun_out/r02_call37.err | head -12; grep -E "warm-up" gpurun_out/r02_call37.err | tail -3
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv
timeout 300 python scratch/rays_probe2.py 2>&1 | head -2 | cut -c1-300
