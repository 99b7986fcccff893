#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): full GPU suite incl. the NCCL test, bench at N=2 (normal and with 4 cores per
# rank, spinning vs blocking host waits), N=1 for reference.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call4.log
: > $OUT
nvidia-smi -L >> $OUT; nproc >> $OUT
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 >> $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "=== N=2" >> $OUT
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err || tail -30 gpurun_out/r02_bench_n2.err >> $OUT
cat gpurun_out/r02_bench_n2.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f M/s' % (j['value']/1e6, j['e2e']['value']/1e6), j['strong_scaling'], j['extra']['setup']['e2e_last_batch'], j['extra']['setup']['blocking_sync'])" >> $OUT 2>&1
echo "=== N=2 on 8 cores, spinning" >> $OUT
CHROMA_B200_SYNC=spin timeout 900 taskset -c 0-7 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_8c_spin.json 2> gpurun_out/r02_bench_n2_8c_spin.err
cat gpurun_out/r02_bench_n2_8c_spin.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f M/s' % (j['value']/1e6, j['e2e']['value']/1e6), j['strong_scaling']['value']/1e6, j['extra']['setup']['e2e_last_batch'], j['extra']['setup']['blocking_sync'])" >> $OUT 2>&1
echo "=== N=2 on 8 cores, auto (blocking)" >> $OUT
timeout 900 taskset -c 0-7 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_8c_block.json 2> gpurun_out/r02_bench_n2_8c_block.err
cat gpurun_out/r02_bench_n2_8c_block.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f M/s' % (j['value']/1e6, j['e2e']['value']/1e6), j['strong_scaling']['value']/1e6, j['extra']['setup']['e2e_last_batch'], j['extra']['setup']['blocking_sync'])" >> $OUT 2>&1
echo "=== N=2 on 4 cores, auto (blocking)" >> $OUT
timeout 900 taskset -c 0-3 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_4c_block.json 2> gpurun_out/r02_bench_n2_4c_block.err
cat gpurun_out/r02_bench_n2_4c_block.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f M/s' % (j['value']/1e6, j['e2e']['value']/1e6), j['strong_scaling']['value']/1e6, j['extra']['setup']['e2e_last_batch'], j['extra']['setup']['blocking_sync'])" >> $OUT 2>&1
echo "=== N=1" >> $OUT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cat gpurun_out/r02_bench_n1.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f M/s' % (j['value']/1e6, j['e2e']['value']/1e6), j['strong_scaling'], j['extra']['setup']['e2e_last_batch'])" >> $OUT 2>&1
cat $OUT
