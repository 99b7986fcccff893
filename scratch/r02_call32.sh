#!/bin/bash
# Round 2, GPU call 19 (8 GPUs): the driver's scaling command at N=8, 4, 2 with the deferred pipeline and the
# single-call upload; host topology of the 8-GPU box.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call32.log
: > $OUT
{ echo "=== topology"; nvidia-smi topo -m | head -12; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)|thread"; nproc;
  for d in /sys/bus/pci/devices/*; do if [ -e $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null && grep -q 0x0302 $d/class 2>/dev/null; then echo "$d numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist)"; fi; done; } >> $OUT 2>&1
for N in ${NS:-8 4 2}; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}_c32.json 2> gpurun_out/r02_bench_n${N}_c32.err || tail -30 gpurun_out/r02_bench_n${N}_c32.err >> $OUT
python -c "
import sys,json
j=json.loads(open('gpurun_out/r02_bench_n${N}_c32.json').read())
print('N=$N value %.1f e2e %.1f M/s ms/step %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step']))
s=j['strong_scaling']; print(' strong %.4f s %.1f M/s %s' % (s['seconds'], s['value']/1e6, s['checksum']))
print(' geometry upload %.1f s blocking %s' % (j['extra']['setup']['upload_geometry_s'], j['extra']['setup']['blocking_sync']))
for r in j['extra']['setup']['per_rank']:
    lb=r['last_batch']
    print(' rank %d max device ms/event %.2f | e2e %.3f s loop %.3f s allreduce+readback %.4f s gap median %.2f max %.2f ms | upload %.2f propagate %.2f total %.2f collect %.2f ms first yield %s ms' % (r['rank'], max(r['device_ms_per_event']), r['e2e_s'], r['loop_s'], r['allreduce_and_readback_s'], r['yield_gap_ms_median'], r['yield_gap_ms_max'], lb['upload_s']*1e3, lb['propagate_s']*1e3, lb['batch_total_s']*1e3, lb.get('collect_s',0)*1e3, r.get('first_yield_ms')))
print(' rank 0 stages', j['extra']['setup']['per_rank'][0].get('stage_log_head'))
" >> $OUT 2>&1
done
cat $OUT
