#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/rays_probe2.py > gpurun_out/r02_call36.log 2>&1; cat gpurun_out/r02_call36.log | cut -c1-420
