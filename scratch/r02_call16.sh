#!/bin/bash
# Round 2, GPU call 16 (1 GPU): host topology, GPU suite, deferred-collection pipeline A/B (bench e2e),
# float-only triangle test vs the literal mixed-precision one (kernel A/B).
mkdir -p gpurun_out
OUT=gpurun_out/r02_call16.log
: > $OUT
{ echo "=== topology"; nvidia-smi topo -m; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)|thread"; nproc;
  for d in /sys/bus/pci/devices/*; do if [ -e $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo "$d numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist)"; fi; done; } >> $OUT 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call16_pytest.log 2>&1; tail -5 gpurun_out/r02_call16_pytest.log
for tag in defer nodefer depth2; do
  case $tag in defer) ENV="";; nodefer) ENV="CHROMA_B200_DEFER=0";; depth2) ENV="CHROMA_B200_PIPELINE_DEPTH=2";; esac
  env $ENV timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_$tag.json 2> gpurun_out/r02_bench_n1_$tag.err
  python - <<P >> $OUT
import json
try:
    j=json.load(open('gpurun_out/r02_bench_n1_$tag.json'))
    r=j['extra']['setup']['per_rank'][0]
    print('$tag value %.1f e2e %.1f M/s ms/step %.3f int0 %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['roofline']['ms_per_launch']), 'last_batch', {k: round(v*1e3,3) if isinstance(v,float) else v for k,v in r['last_batch'].items()}, 'gap', r['yield_gap_ms_median'], r['yield_gap_ms_max'], 'strong', j['strong_scaling']['seconds'], j['strong_scaling']['checksum'])
except Exception as e:
    print('$tag failed', e)
P
done
echo "=== kernel A/B: float triangle test (default) vs literal double" >> $OUT
timeout 600 python scratch/sweep.py "TRACE=1" "" >> $OUT 2>&1
CHROMA_B200_LIB=$PWD/scratch/lib_tridouble.so timeout 600 python scratch/sweep.py "TRACE=1" "" >> $OUT 2>&1
grep -vE "^\s+all" $OUT | cut -c1-400
