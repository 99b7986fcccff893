#!/bin/bash
# Round 2, GPU call 28 (8 GPUs): what bounds the end-to-end rate at 8 ranks -- upload variants in one launch.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 400 $TR scratch/e2e_probe8.py > gpurun_out/r02_call28.log 2> gpurun_out/r02_call28.err || tail -20 gpurun_out/r02_call28.err
cat gpurun_out/r02_call28.log
