#!/bin/bash
# Round 2, GPU call 17 (1 GPU): upload probe (write-combined, threads, busy GPU), propagate host overhead trimmed.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call17.log
: > $OUT
timeout 600 python scratch/h2d_probe2.py >> $OUT 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_c17.json 2> gpurun_out/r02_bench_n1_c17.err
python - <<P >> $OUT
import json
j=json.load(open('gpurun_out/r02_bench_n1_c17.json'))
r=j['extra']['setup']['per_rank'][0]
print('value %.1f e2e %.1f M/s ms/step %.3f int0 %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['roofline']['ms_per_launch']), 'last_batch', {k: round(v*1e3,3) if isinstance(v,float) else v for k,v in r['last_batch'].items()}, 'strong', j['strong_scaling']['seconds'], j['strong_scaling']['checksum'])
P
timeout 600 python -m pytest tests -m gpu -q -x -k "propagate or fixtures or full_size" 2>&1 | tail -3 >> $OUT
cat $OUT | cut -c1-500
