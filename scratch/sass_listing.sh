#!/bin/bash
# SASS listings of the hot kernels + a census per kernel -> profiles/<tag>_sass_<kernel>.txt, <tag>_sass_census.txt
tag=${1:-r02}
lib=chroma_lite_b200/libchroma_b200.so
: > profiles/${tag}_sass_census.txt
for k in step_intersect_kernelILb0E propagate_tail_lanes_kernelILb0ELb0E propagate_tail_kernelILb0ELb0E step_physics_kernelILb0E; do
  sym=$(cuobjdump -elf $lib 2>/dev/null | grep -o "_ZN2cb[0-9]*${k}[A-Za-z0-9_]*" | sort -u | head -1)
  short=$(echo $k | sed 's/ILb.*//')
  cuobjdump -sass -fun "$sym" $lib 2>/dev/null | grep -vE "^\s*$|cuobjdump warning" | sed -E "s#/\* 0x[0-9a-f]* \*/##; s/[[:space:]]+/ /g" > profiles/${tag}_sass_${short}.txt
  {
    echo "== $short ($sym)"
    echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' profiles/${tag}_sass_${short}.txt)"
    awk '$1 ~ /^\/\*[0-9a-f]+\*\/$/ {print $2}' profiles/${tag}_sass_${short}.txt | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -24 | awk '{printf "  %s=%s", $2, $1} END {print ""}'
    echo "  local memory: LDL=$(grep -c 'LDL' profiles/${tag}_sass_${short}.txt) STL=$(grep -c 'STL' profiles/${tag}_sass_${short}.txt)  bulk copy (TMA 1-D): UBLKCP=$(grep -c UBLKCP profiles/${tag}_sass_${short}.txt)  warp reductions: REDUX=$(grep -c REDUX profiles/${tag}_sass_${short}.txt)  MATCH=$(grep -c 'MATCH' profiles/${tag}_sass_${short}.txt)  PRMT=$(grep -c PRMT profiles/${tag}_sass_${short}.txt) DFMA=$(grep -c DFMA profiles/${tag}_sass_${short}.txt) CCTL/prefetch=$(grep -c 'CCTL' profiles/${tag}_sass_${short}.txt)"
  } >> profiles/${tag}_sass_census.txt
done
cat profiles/${tag}_sass_census.txt
