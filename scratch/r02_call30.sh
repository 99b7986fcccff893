#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/e2e_probe8.py > gpurun_out/r02_call30.log 2>&1; cat gpurun_out/r02_call30.log | cut -c1-400
