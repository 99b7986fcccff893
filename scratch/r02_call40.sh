#!/bin/bash
# Round 2, GPU call 40 (1 GPU): the same bench line over 100 events (pipeline fill and drain amortised).
mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 5 --cpu-sample 200000 > gpurun_out/r02k_bench_100steps.json 2> gpurun_out/r02k_bench_100steps.err || tail -20 gpurun_out/r02k_bench_100steps.err
python -c "
import json, numpy as np
j=json.load(open('gpurun_out/r02k_bench_100steps.json')); r=j['extra']['setup']['per_rank'][0]
ms=np.array(r['device_ms_per_event'])
print('100 events: value %.1f e2e %.1f M/s ms/step %.3f | device ms/event min %.2f median %.2f p95 %.2f max %.2f | gap median %.2f ms' % (j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], ms.min(), np.median(ms), np.percentile(ms,95), ms.max(), r['yield_gap_ms_median']))"
