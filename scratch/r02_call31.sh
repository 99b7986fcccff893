#!/bin/bash
mkdir -p gpurun_out
VARIANTS=2 CHROMA_B200_TRACE=1 timeout 600 python scratch/e2e_probe8.py > gpurun_out/r02_call31.log 2>&1
grep -E "gpu stages|e2e " gpurun_out/r02_call31.log | cut -c1-700
grep -nE "tail: [0-9]+ photons [0-9]{2,}\.[0-9]+ ms|intersect [0-9]{2,}\.[0-9]+ ms" gpurun_out/r02_call31.log | head
