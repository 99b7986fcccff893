#!/bin/bash
# Round-2 A/B of the traversal-tree options on one B200 (each variant needs its own process: the tree
# is built in cb_geometry_create).  usage: gpurun --timeout 1500 -- 'bash scratch/ab_tree.sh'
# Output: gpurun_out/ab_tree.log (one line per variant: photons/s, e2e, first-step traversal ms,
# geometry upload incl. tree build).  Per-step times: add CHROMA_B200_TRACE=1 and use scratch/sweep.py.
mkdir -p gpurun_out
OUT=gpurun_out/ab_tree.log
: > $OUT
run() {
    name=$1; shift
    env "$@" python bench.py --steps 8 --warmup 3 --cpu-sample 20000 2>gpurun_out/ab_tree_$name.err | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('%-22s value %.1f M/s  e2e %.1f M/s  first-step traversal %.3f ms  event %.3f ms  geometry upload %.1f s' % (
    '$name', j['value'] / 1e6, j['e2e']['value'] / 1e6, j['roofline']['ms_per_launch'], j['ms_per_step'],
    j['extra']['setup']['upload_geometry_s']))" >> $OUT
}
run default            X=1
run single             CHROMA_B200_TREE=single
run split8             CHROMA_B200_LEAF_SPLIT=8,8,2
run single_split8      CHROMA_B200_TREE=single CHROMA_B200_LEAF_SPLIT=8,8,2
run single_split4      CHROMA_B200_TREE=single CHROMA_B200_LEAF_SPLIT=4,8,2
# single level holds more stack entries (scratch/emu_stack.py: 1.0 % of the expansions above 16 vs 0.05 %): same tree with
# 24 shared-memory entries per lane; build the variant first with: bash scratch/mkvariant.sh pstack24 "-DCB_PSTACK_N=24"
[ -f scratch/lib_pstack24.so ] && run single_pstack24 CHROMA_B200_TREE=single CHROMA_B200_LIB=$PWD/scratch/lib_pstack24.so
run default_again      X=1
cat $OUT
