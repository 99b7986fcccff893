#!/bin/bash
# Round 2, GPU calls 5 and 8 (8 GPUs): the driver's scaling command at N=8, host waits auto (spin-then-block) vs spinning.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call8.log
nvidia-smi --query-gpu=index,pci.bus_id,power.limit,clocks.max.sm,temperature.gpu --format=csv >> /dev/null
: > $OUT
nvidia-smi --query-gpu=index,pci.bus_id,power.limit,clocks.max.sm,temperature.gpu --format=csv >> $OUT
nvidia-smi -L | wc -l >> $OUT; nproc >> $OUT; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))" >> $OUT
cat /sys/fs/cgroup/cpu.max >> $OUT 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
summ() { python -c "
import sys,json
j=json.loads(open('$1').read())
print('value %.1f e2e %.1f M/s strong %.1f M/s checksum %s blocking %s' % (j['value']/1e6, j['e2e']['value']/1e6, j['strong_scaling']['value']/1e6, j['strong_scaling']['checksum'], j['extra']['setup']['blocking_sync']))
print(' geometry upload %.1f s' % j['extra']['setup']['upload_geometry_s'])
for r in j['extra']['setup']['per_rank']:
    lb=r['last_batch']
    if 'device_ms_per_event' in r: print(' rank %d %s device ms/event: %s' % (r['rank'], r.get('pci_bus_id'), ' '.join('%.2f' % x for x in r['device_ms_per_event'])))
    print(' rank %d e2e %.3f s loop %.3f s allreduce+readback %.4f s gap median %.2f max %.2f ms | upload %.2f propagate %.2f readback %.2f daq %.2f ms cores %d' % (r['rank'], r['e2e_s'], r['loop_s'], r['allreduce_and_readback_s'], r['yield_gap_ms_median'], r['yield_gap_ms_max'], lb['upload_s']*1e3, lb['propagate_s']*1e3, lb['readback_s']*1e3, lb['daq_s']*1e3, r['affinity_cores']))
" >> $OUT 2>&1; }
echo "=== N=8 auto" >> $OUT
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err || tail -30 gpurun_out/r02_bench_n8.err >> $OUT
summ gpurun_out/r02_bench_n8.json
echo "=== N=8 spin, no clock sampler" >> $OUT
CHROMA_B200_NO_CLOCKS=1 CHROMA_B200_SYNC=spin timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_spin.json 2> gpurun_out/r02_bench_n8_spin.err || tail -30 gpurun_out/r02_bench_n8_spin.err >> $OUT
summ gpurun_out/r02_bench_n8_spin.json
cat $OUT
