#!/bin/bash
# Round 2, GPU call 18 (1 GPU): single-call bank upload, kernel-setup cache fix; GPU suite + bench.
mkdir -p gpurun_out
OUT=gpurun_out/r02_call18.log
: > $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call18_pytest.log 2>&1; tail -15 gpurun_out/r02_call18_pytest.log | cut -c1-300 >> $OUT
for tag in a b; do
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_c18$tag.json 2> gpurun_out/r02_bench_n1_c18$tag.err
python - <<P >> $OUT
import json
j=json.load(open('gpurun_out/r02_bench_n1_c18$tag.json'))
r=j['extra']['setup']['per_rank'][0]
print('value %.1f e2e %.1f M/s ms/step %.3f int0 %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['roofline']['ms_per_launch']), 'last_batch', {k: round(v*1e3,3) if isinstance(v,float) else v for k,v in r['last_batch'].items()}, 'gap', r['yield_gap_ms_median'], 'strong', j['strong_scaling']['seconds'], j['strong_scaling']['checksum'])
P
done
timeout 300 python scratch/h2d_probe2.py 2>&1 | grep "write_combined=False" >> $OUT
cat $OUT | cut -c1-600
