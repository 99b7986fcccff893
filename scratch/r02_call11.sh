#!/bin/bash
# Round 2, GPU calls 11 and 15 (8 GPUs): the driver's scaling command at N=8 after the untraceable-ray fix.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call15.log
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_v2.json 2> gpurun_out/r02_bench_n8_v2.err || tail -30 gpurun_out/r02_bench_n8_v2.err >> $OUT
python -c "
import sys,json
j=json.loads(open('gpurun_out/r02_bench_n8_v2.json').read())
print('value %.1f e2e %.1f M/s ms/step %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step']))
print(j['strong_scaling'])
print(' geometry upload %.1f s blocking %s' % (j['extra']['setup']['upload_geometry_s'], j['extra']['setup']['blocking_sync']))
for r in j['extra']['setup']['per_rank']:
    lb=r['last_batch']
    print(' rank %d max device ms/event %.2f | e2e %.3f s loop %.3f s allreduce+readback %.4f s gap median %.2f max %.2f ms | upload %.2f propagate %.2f readback %.2f daq %.2f ms' % (r['rank'], max(r['device_ms_per_event']), r['e2e_s'], r['loop_s'], r['allreduce_and_readback_s'], r['yield_gap_ms_median'], r['yield_gap_ms_max'], lb['upload_s']*1e3, lb['propagate_s']*1e3, lb['readback_s']*1e3, lb['daq_s']*1e3))
" >> $OUT 2>&1
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532"
timeout 900 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4_v2.json 2> gpurun_out/r02_bench_n4_v2.err
python -c "
import json; j=json.load(open('gpurun_out/r02_bench_n4_v2.json')); print('N=4 value %.1f e2e %.1f M/s' % (j['value']/1e6, j['e2e']['value']/1e6), j['strong_scaling']['value']/1e6, j['strong_scaling']['checksum'])" >> $OUT 2>&1
cat $OUT
