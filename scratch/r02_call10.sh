#!/bin/bash
# Round 2, GPU call 10 (1 GPU): the bench lines of the other configs, both arms: rays (config 2, lion x16),
# scint (config 4), pdf (f-2), then the heavy 29k-PMT variant (169.8 M triangles) when the host has the memory.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call10.log
: > $OUT
free -g | head -2 >> $OUT; df -h /tmp | tail -1 >> $OUT
for w in rays scint pdf; do
  for impl in reference ours; do
    timeout 900 python bench.py --workload $w --impl $impl > gpurun_out/r02_bench_${w}_${impl}.json 2> gpurun_out/r02_bench_${w}_${impl}.err
    echo "== $w $impl rc=$?" >> $OUT
    python -c "
import json; j=json.load(open('gpurun_out/r02_bench_${w}_${impl}.json')); print('   value %.4g %s  ms/step %.3f  e2e %.4g' % (j['value'], j['unit'], j['ms_per_step'], j['e2e']['value']))" >> $OUT 2>&1
  done
done
mem=$(free -g | awk '/Mem:/{print $7}')
echo "available host memory: $mem GB" >> $OUT
if [ "$mem" -gt 150 ]; then
  CHROMA_B200_CACHE=/dev/shm/cb_cache_heavy timeout 1500 python bench.py --workload pmt29k_heavy --steps 5 --warmup 3 --cpu-sample 100000 > gpurun_out/r02_bench_heavy_ours.json 2> gpurun_out/r02_bench_heavy_ours.err
  echo "== heavy ours rc=$?" >> $OUT
  tail -3 gpurun_out/r02_bench_heavy_ours.err >> $OUT
  python -c "
import json; j=json.load(open('gpurun_out/r02_bench_heavy_ours.json')); print('   value %.4g %s  ms/step %.3f  e2e %.4g triangles %d' % (j['value'], j['unit'], j['ms_per_step'], j['e2e']['value'], j['config']['triangles'])); print(j['extra']['setup'])" >> $OUT 2>&1
  rm -rf /dev/shm/cb_cache_heavy
fi
cat $OUT
