#!/bin/bash
# A/B timing of library variants: prints value/ms per variant
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset CHROMA_B200_LIB; else export CHROMA_B200_LIB=/root/repo/scratch/$lib; fi
  timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('LIB=$lib value %.4g ms/step %.3f e2e %.4g launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))"
done
